#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
B200MP_SIFT=1 timeout 200 python tools/states_probe.py > $O/states.log 2>&1
B200MP_SIFT=0 timeout 200 python tools/states_probe.py >> $O/states.log 2>&1
cat $O/states.log
