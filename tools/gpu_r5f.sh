#!/bin/bash
# sifting pass depth (distal links tested before the survivors are packed): roadmap sample phase for 2 / 3 / 4 / 6 (default) / 8 links
O=gpurun_out/$1
mkdir -p $O
timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap_p6.log 2>&1
for v in p2 p3 p4 p8; do B200MP_LIB=$PWD/ompl_b200/variants/libb200mp_$v.so timeout 300 python tools/roadmap_probe.py 1048576 > $O/roadmap_$v.log 2>&1; done
tail -n 1 $O/roadmap_*.log | cut -c1-260
