#!/bin/bash
# chunked host path with two alternating kernel streams: parity of the chunked entry points, e2e with one / two streams (three runs each)
O=gpurun_out/$1
mkdir -p $O
timeout 600 python -m pytest tests/test_validate_gpu.py -m gpu -x -q -k "chunked or linearity or c2_sample or first_look or double_precision" 2>&1 | tail -5 > $O/pytest.log
for i in 1 2 3; do
  echo "two streams: $(timeout 200 python tools/e2e_probe.py 2>&1 | grep -E 'f32|f64' | tr '\n' ' ')"
  echo "one stream : $(B200MP_CHUNK_STREAMS=1 timeout 200 python tools/e2e_probe.py 2>&1 | grep -E 'f32|f64' | tr '\n' ' ')"
done > $O/e2e.log 2>&1
cat $O/pytest.log $O/e2e.log
