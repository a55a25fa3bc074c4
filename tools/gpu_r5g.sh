#!/bin/bash
# plan constants in shared memory + roadmap-edge first look at 3/8 and 6/8: parity, C2, roadmap (twice), smoke
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
timeout 200 python tools/r2_probe.py gridonly > $O/c2.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
cat $O/pytest.log; tail -n 1 $O/c2.log $O/roadmap.log $O/roadmap2.log $O/smoke.log
