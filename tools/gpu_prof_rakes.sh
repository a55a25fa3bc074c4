#!/bin/bash
# one full ncu capture of the packed-rakes kernel on roadmap edges (launch index as in tools/gpu_round2.sh)
O=gpurun_out/$1
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:validate_rakes -s 8 -c 1 -o $O/prof_rakes_roadmap python tools/profile_target.py > $O/prof.log 2>&1
tail -3 $O/prof.log; ls -la $O
