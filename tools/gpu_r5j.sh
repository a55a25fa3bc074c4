#!/bin/bash
# bench line + reference arm only (profiles refresh after a host-path change)
O=gpurun_out/$1
mkdir -p $O
timeout 170 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench_n1.err
cut -c1-700 $O/bench_n1.json; cut -c1-200 $O/bench_reference.json; tail -2 $O/bench_n1.err
