#!/bin/bash
# e2e time of b200mp_validate_motions (pinned host buffers, 2^20 C2 motions) for a few chunk plans of the host pipeline
O=gpurun_out/$1
mkdir -p $O
for plan in "262144 524288" "131072 524288" "131072 262144" "262144 262144" "65536 524288" "524288 524288" "1048576 1048576"; do
  set -- $plan
  echo "chunk0=$1 chunkmax=$2: $(B200MP_CHUNK0=$1 B200MP_CHUNKMAX=$2 timeout 200 python tools/e2e_probe.py 2>&1 | grep f32)"
done | tee $O/e2e_sweep.log
