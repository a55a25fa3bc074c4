#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_roadmap_gpu.py tests/test_validate_gpu.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest.log
B200MP_EDGES_PACKED=0 timeout 300 python tools/roadmap_probe.py 262144 1048576 > $O/roadmap_unpacked.log 2>&1
timeout 300 python tools/roadmap_probe.py 262144 1048576 > $O/roadmap_packed.log 2>&1
cat $O/pytest.log $O/roadmap_unpacked.log $O/roadmap_packed.log
