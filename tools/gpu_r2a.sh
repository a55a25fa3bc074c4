#!/bin/bash
# round 2, first session: gpu tests, probes, bench with the roadmap parity block
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/pytest.log
timeout 600 python tools/r2_probe.py > $O/probe.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/pytest.log; cat $O/probe.log; cat $O/bench.json | tail -c 2500; tail -3 $O/bench.err
