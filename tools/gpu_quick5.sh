#!/bin/bash
# whole GPU suite + C2 variants + roadmap + C4 + session probe
O=gpurun_out/$1
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest.log
timeout 300 python tools/r2_probe.py dense > $O/dense.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
timeout 300 python tools/configs_probe.py --only c4 > $O/c4.log 2>&1
timeout 120 python tools/session_probe.py > $O/session_probe.log 2>&1
cat $O/pytest.log $O/dense.log $O/roadmap.log $O/session_probe.log; tail -1 $O/c4.log | cut -c1-800
