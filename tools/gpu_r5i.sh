#!/bin/bash
# chunk plans of the host path with two kernel streams
O=gpurun_out/$1
mkdir -p $O
for plan in "262144 524288" "262144 262144" "131072 262144" "131072 524288" "196608 393216"; do
  set -- $plan
  echo "chunk0=$1 chunkmax=$2: $(B200MP_CHUNK0=$1 B200MP_CHUNKMAX=$2 timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)"
done > $O/e2e_sweep.log 2>&1
cat $O/e2e_sweep.log
