#!/bin/bash
# quick loop + C4 probe + one ncu capture of the packed-rakes kernel on roadmap edges
O=gpurun_out/$1
mkdir -p $O
timeout 900 python -m pytest tests/test_validate_gpu.py tests/test_roadmap_gpu.py tests/test_capt_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
timeout 300 python tools/r2_probe.py dense > $O/dense.log 2>&1
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap.log 2>&1
timeout 300 python tools/configs_probe.py --only c4 > $O/c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:validate_rakes -s 8 -c 1 -o $O/prof_rakes_roadmap python tools/profile_target.py > $O/prof.log 2>&1
cat $O/pytest.log $O/dense.log $O/roadmap.log; tail -3 $O/c4.log | cut -c1-700
