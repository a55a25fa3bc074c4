#!/bin/bash
# registers / spills / stack of every kernel of one translation unit: tools/ptxas_report.sh validate_panda.cu
cd "$(dirname "$0")/../ompl_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -c "$1" -o /tmp/ptxas_report.o 2>&1 \
 | awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /Function properties for/ {fn=$0; sub(/.*for /,"",fn)} /bytes stack frame/ {st[fn]=$0} /Used [0-9]+ registers/ {print name; print "   " st[name]; print "   " $0}' | c++filt | cut -c1-220
