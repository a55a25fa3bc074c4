#!/bin/bash
# quick loop: validation + roadmap + capt tests, then the C2 variants and the 2^20 roadmap phases
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py tests/test_capt_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 300 python tools/r2_probe.py dense > $O/dense.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
cat $O/pytest.log $O/dense.log $O/roadmap.log
