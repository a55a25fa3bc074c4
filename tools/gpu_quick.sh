#!/bin/bash
# Short GPU-box session while iterating on a kernel: parity tests, the dev check (parity + rough throughput), a short bench.
O=gpurun_out/$1
mkdir -p $O
python -m pytest tests -m gpu -x -q ${2:-} 2>&1 | tail -15 > $O/pytest.log
python tools/dev_gpu_check.py > $O/devcheck.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
cat $O/pytest.log; cat $O/devcheck.log; cut -c1-400 $O/bench.json; tail -3 $O/bench.err
