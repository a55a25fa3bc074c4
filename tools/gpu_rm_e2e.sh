#!/bin/bash
# roadmap tests + the bench's roadmap block (build_e2e_ms = wall clock incl. the copy of the roadmap to host memory)
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_roadmap_gpu.py tests/test_adapters.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-configs > $O/bench.json 2> $O/bench.err
cat $O/pytest.log; tail -c 1800 $O/bench.json; tail -3 $O/bench.err
