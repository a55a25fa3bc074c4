#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
bash tools/gpu_flops.sh $1 > /dev/null 2>&1
cp $O/validate_flops.json profiles/validate_flops.json
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json,sys
d=json.load(open('$O/bench_n1.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','roadmap_build_ms','roofline','c4','c5')}, indent=None)[:6000])
print(json.dumps(d['roadmap'])[:3000])
"
tail -3 $O/bench_n1.err
