#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest.log
B200MP_MOTIONS_BUCKETED=0 timeout 300 python tools/r2_probe.py dense 2>&1 | head -1 > $O/c2_plain.log
B200MP_FIRST_LOOK=0 timeout 300 python tools/r2_probe.py dense 2>&1 | head -1 > $O/c2_bucketed_nofirst.log
timeout 300 python tools/r2_probe.py dense 2>&1 | head -1 > $O/c2_firstlook.log
cat $O/pytest.log $O/c2_plain.log $O/c2_bucketed_nofirst.log $O/c2_firstlook.log
