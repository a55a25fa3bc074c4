#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 300 python -m pytest tests/test_session_gpu.py tests/test_c1_rrtconnect_gpu.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest_session.log
timeout 120 python tools/session_probe.py > $O/session_probe.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
cat $O/pytest_session.log $O/session_probe.log $O/roadmap.log
