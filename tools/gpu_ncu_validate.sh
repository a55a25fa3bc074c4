#!/bin/bash
# ncu --set full of the validation kernel on the bench workloads (launch #1 = C2 mixed 2^20, #2 = C2 all-valid 2^17)
O=gpurun_out/$1
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:validate_motions -s 1 -c 2 -o $O/prof_validate python tools/profile_target.py --no-roadmap > $O/target.log 2>&1
ls -la $O
