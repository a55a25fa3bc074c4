#!/bin/bash
# 2-GPU session: all gpu tests (incl. the 2-GPU ones), bench at N=1 and N=2
O=gpurun_out/$1
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
cat $O/pytest.log; tail -c 1500 $O/bench_n1.json; tail -3 $O/bench_n1.err; tail -c 1500 $O/bench_n2.json; tail -5 $O/bench_n2.err
