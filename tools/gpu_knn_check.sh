#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_roadmap_gpu.py tests/test_chain_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
timeout 300 python tools/r2_probe.py knn > $O/knn.log 2>&1
cat $O/pytest.log $O/roadmap.log; head -4 $O/knn.log
