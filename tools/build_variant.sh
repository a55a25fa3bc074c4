#!/bin/bash
# Development: a second libb200mp.so whose Panda validation kernels are compiled with extra -D flags, for A/B runs on the GPU box
# (B200MP_LIB=<path> selects it in ompl_b200/engine.py). usage: tools/build_variant.sh <tag> -DNAME=VALUE ...
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p build/obj ompl_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c ompl_b200/csrc/validate_panda.cu -o build/obj/validate_panda_$tag.o
objs=$(ls build/obj/*.o | grep -v "validate_panda")
nvcc -shared -o ompl_b200/variants/libb200mp_$tag.so $objs build/obj/validate_panda_$tag.o
ls -la ompl_b200/variants/libb200mp_$tag.so
