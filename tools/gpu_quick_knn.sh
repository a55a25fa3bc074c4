#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
python -m pytest tests/test_knn_gpu.py tests/test_roadmap_gpu.py tests/test_chain_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
cat $O/pytest.log; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], json.dumps(d['roadmap']))"; tail -3 $O/bench.err
