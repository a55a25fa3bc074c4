#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest.log
cat $O/pytest.log
