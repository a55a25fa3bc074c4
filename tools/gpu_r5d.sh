#!/bin/bash
# first look that gives up at QUIT undecided lanes: parity (default build), C2 for QUIT = 0 / 2 / 4 / 8 / 12, roadmap, e2e
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
timeout 200 python tools/r2_probe.py gridonly > $O/c2_q4.log 2>&1
for v in q0 q2 q8 q12; do B200MP_LIB=$PWD/ompl_b200/variants/libb200mp_$v.so timeout 200 python tools/r2_probe.py gridonly > $O/c2_$v.log 2>&1; done
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
echo "e2e: $(timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)" > $O/e2e.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $O/launches_c2.csv python tools/r2_probe.py gridonly > /dev/null 2>&1
cat $O/pytest.log; tail -n 1 $O/c2_*.log $O/roadmap.log; cat $O/e2e.log
