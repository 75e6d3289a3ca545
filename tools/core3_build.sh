#!/bin/bash
# Dev aid: build variants of the one-sample-per-step cluster kernel (H=192 instantiation) as separate libraries under lib/ab/.
#   tools/core3_build.sh name "-DSWN_DBG_HELP_SLEEP_NS=8000" [...]  ->  lib/ab/libswn_core3_<name>.so  (select with SWN_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../shallow-wavenet_b200/csrc"
OBJ=../lib/obj; AB=../lib/ab; mkdir -p $AB
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON=$(ls $OBJ/*.o | grep -v swn_core_nch3.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DSWN_NCH=3 $flags -c swn_core_inst.cu -o $AB/core3_$name.o \
    && nvcc $ARCH -shared -o $AB/libswn_core3_$name.so $COMMON $AB/core3_$name.o -cudart static && echo built $name ) &
done
wait
