#!/bin/bash
# long kernel's chunked takes + bucket plan without bank conflicts: parity, C2 with 1 / 2 / 4 first-look lanes, roadmap, e2e chunk plans
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 200 python tools/r2_probe.py gridonly > $O/c2_fl2.log 2>&1
for v in fl1 fl4; do B200MP_LIB=$PWD/ompl_b200/variants/libb200mp_$v.so timeout 200 python tools/r2_probe.py gridonly > $O/c2_$v.log 2>&1; done
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
for plan in "262144 524288" "262144 262144" "131072 262144" "131072 131072"; do
  set -- $plan
  echo "chunk0=$1 chunkmax=$2: $(B200MP_CHUNK0=$1 B200MP_CHUNKMAX=$2 timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)"
done > $O/e2e_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_c2.csv python tools/r2_probe.py gridonly > /dev/null 2>&1
cat $O/pytest.log; tail -n 2 $O/c2_*.log; tail -n 1 $O/roadmap.log; cat $O/e2e_sweep.log; grep -c . $O/launches_c2.csv
