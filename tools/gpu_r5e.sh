#!/bin/bash
# pre-pass with 1024-motion blocks + coalesced plan kernel: parity, C2, roadmap, e2e, launch list of the C2 step
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py tests/test_capt_gpu.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
timeout 200 python tools/r2_probe.py gridonly > $O/c2.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
echo "e2e: $(timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)" > $O/e2e.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $O/launches_c2.csv python tools/r2_probe.py gridonly > /dev/null 2>&1
cat $O/pytest.log; tail -n 1 $O/c2.log $O/roadmap.log; cat $O/e2e.log; grep "bucket" $O/launches_c2.csv | head -6 | cut -d, -f5,12-
