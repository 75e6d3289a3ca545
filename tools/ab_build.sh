#!/bin/bash
# Dev aid: build A/B variants of the v2 cluster kernel (H=256 instantiation only) as separate libraries.
#   tools/ab_build.sh name "-DSWN_DEFER_BG=0" [name2 "flags2" ...]   ->  shallow-wavenet_b200/lib/ab/core4_<name>.o (linked on the GPU box by tools/ab_run.py)
# Select one at run time with SWN_B200_LIB=<path>; tools/ab_run.py times them against each other.
set -e
cd "$(dirname "$0")/../shallow-wavenet_b200/csrc"
OBJ=../lib/obj; AB=../lib/ab; mkdir -p $AB
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DSWN_NCH=4 $flags -c swn_core_inst.cu -o $AB/core4_$name.o 2> $AB/core4_$name.log \
    && grep -A1 "swn_core_kernelILi4ELi1" $AB/core4_$name.log | tail -1 | sed "s/^/$name: /" ) &
done
wait
