#!/bin/bash
# one multi-GPU bench run as the driver launches it: tools/gpu_scale_one.sh <tag> <N> [extra bench flags]
O=gpurun_out/$1
N=$2
shift 2
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 5 --warmup 3 "$@" > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 1600 $O/bench_n$N.json; tail -3 $O/bench_n$N.err
