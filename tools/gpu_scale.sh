#!/bin/bash
# Multi-GPU session (gpurun --gpus 8): the bench at N = 8, 4, 2 as the driver launches it, then the roadmap probe on 8 GPUs.
O=gpurun_out/$1
mkdir -p $O
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 tools/roadmap_probe_mgpu.py > $O/roadmap_probe_8gpu.log 2>&1
for N in 8 4 2; do cut -c1-300 $O/bench_n$N.json; tail -2 $O/bench_n$N.err; done; grep "^0 2" $O/roadmap_probe_8gpu.log
