#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/pytest.log
timeout 300 python tools/roadmap_probe.py 262144 1048576 > $O/roadmap.log 2>&1
timeout 300 python tools/r2_probe.py dense > $O/dense.log 2>&1
cat $O/pytest.log $O/roadmap.log $O/dense.log
