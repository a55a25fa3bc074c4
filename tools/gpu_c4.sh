#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 600 python -m pytest tests/test_capt_gpu.py tests/test_filter_gpu.py -m gpu -x -q 2>&1 | tail -3 > $O/pytest.log
timeout 300 python tools/configs_probe.py --only c4 > $O/c4.log 2>&1
cat $O/pytest.log; tail -1 $O/c4.log | cut -c1-900
