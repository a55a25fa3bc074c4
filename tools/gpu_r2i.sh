#!/bin/bash
O=gpurun_out/$1
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:validate_rakes -c 1 -o $O/prof_rakes python tools/profile_target.py > $O/t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:validate_motions -s 2 -c 1 -o $O/prof_motions_allvalid python tools/profile_target.py --no-roadmap > $O/t2.log 2>&1
ls -la $O
