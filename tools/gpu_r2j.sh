#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_adapters.py -m gpu -x -q 2>&1 | tail -30 > $O/pytest_adapters.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest.log
cat $O/pytest_adapters.log $O/pytest.log
