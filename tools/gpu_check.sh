#!/bin/bash
# checkpoint: whole GPU test suite + the bench line (with the roadmap parity block against the compiled reference)
O=gpurun_out/$1
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
cat $O/pytest.log; tail -c 2500 $O/bench_n1.json; tail -5 $O/bench_n1.err
