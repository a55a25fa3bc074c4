#!/bin/bash
# attachments: parity tests + the whole GPU suite + perf probes (does the attachment branch cost the hot kernels anything?)
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py -m gpu -x -q -k "attach" 2>&1 | tail -8 > $O/pytest_attach.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log
timeout 300 python tools/r2_probe.py dense > $O/dense.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
cat $O/pytest_attach.log $O/pytest.log $O/dense.log $O/roadmap.log
