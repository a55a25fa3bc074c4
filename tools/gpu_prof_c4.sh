#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:validate_motions -s 1 -c 1 -o $O/prof_c4 python tools/profile_c4.py > $O/prof_c4.log 2>&1
tail -2 $O/prof_c4.log
