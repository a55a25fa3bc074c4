#!/bin/bash
# first look through the state pipeline: parity, C2, roadmap, e2e, launch list of the C2 step
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q 2>&1 | tail -15 > $O/pytest.log
timeout 200 python tools/r2_probe.py gridonly > $O/c2.log 2>&1
B200MP_LOOK_STATES=0 timeout 200 python tools/r2_probe.py gridonly > $O/c2_old.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
B200MP_LOOK_STATES=0 timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap_old.log 2>&1
for plan in "262144 524288" "262144 262144" "131072 524288"; do
  set -- $plan
  echo "chunk0=$1 chunkmax=$2: $(B200MP_CHUNK0=$1 B200MP_CHUNKMAX=$2 timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)"
done > $O/e2e_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_c2.csv python tools/r2_probe.py gridonly > /dev/null 2>&1
cat $O/pytest.log; tail -n 1 $O/c2.log $O/c2_old.log $O/roadmap.log $O/roadmap_old.log; cat $O/e2e_sweep.log
