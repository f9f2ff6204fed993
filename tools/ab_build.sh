#!/bin/bash
# A/B builds of liblrpt_b200.so with extra -D flags on ONE kernel source (default demod_lane.cu):
#   [SRC=demod_ws.cu] tools/ab_build.sh <name> <flags...>
# -> meteor_demod_b200/ab_<name>.so (load with LRPT_SO=...). Needs the normal build's objects.
set -e
cd "$(dirname "$0")/../meteor_demod_b200/csrc"
src=${SRC:-demod_lane.cu}
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off \
     -I ../../include -I . "$@" -c $src -o _obj/ab_$name.o 2>&1 | grep -E "error" || true
objs=$(ls _obj/*.o | grep -v "_obj/ab_" | grep -v "_obj/$src.o")
nvcc -shared -o ../ab_$name.so $objs _obj/ab_$name.o -cudart static -lm 2>&1 | grep -v deprecat || true
ls -la ../ab_$name.so
