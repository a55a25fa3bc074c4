#!/bin/bash
# the resident single-call kernel: tests (bounded by timeouts: a resident kernel must never be able to hold the box), latency probe, e2e sweep
O=gpurun_out/$1
mkdir -p $O
timeout 300 python -m pytest tests/test_session_gpu.py tests/test_c1_rrtconnect_gpu.py -m gpu -x -q 2>&1 | tail -12 > $O/pytest_session.log
timeout 120 python tools/session_probe.py > $O/session_probe.log 2>&1
timeout 300 python -m pytest tests/test_adapters.py -m gpu -x -q 2>&1 | tail -5 > $O/pytest_adapters.log
cat $O/pytest_session.log $O/session_probe.log $O/pytest_adapters.log
bash tools/gpu_e2e_sweep.sh $1
