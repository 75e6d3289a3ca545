#!/bin/bash
# Dev aid: build variants of the seg cluster kernel (H=192 instantiation) as separate libraries under shallow-wavenet_b200/lib/ab/.
#   tools/seg_build.sh name "-DSWN_STAMPS=1" [name2 "flags2" ...]  ->  lib/ab/libswn_seg_<name>.so  (select with SWN_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../shallow-wavenet_b200/csrc"
OBJ=../lib/obj; AB=../lib/ab; mkdir -p $AB
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON=$(ls $OBJ/*.o | grep -v swn_coreseg_nch3.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DSWN_NCH=3 $flags -c swn_coreseg_inst.cu -o $AB/coreseg3_$name.o 2> $AB/coreseg3_$name.log \
    && nvcc $ARCH -shared -o $AB/libswn_seg_$name.so $COMMON $AB/coreseg3_$name.o -cudart static \
    && grep -A2 "coreseg_kernelILi3ELi5" $AB/coreseg3_$name.log | grep "registers\|spill" | tr '\n' ' ' | sed "s/^/$name: /" && echo ) &
done
wait
