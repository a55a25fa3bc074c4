#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 5 --warmup 3 --no-parity > $O/bench_n$N.json 2> $O/bench_n$N.err
  tail -c 900 $O/bench_n$N.json; echo; tail -2 $O/bench_n$N.err
done
