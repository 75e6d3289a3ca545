#!/bin/bash
# Dev aid: diagnostic / A-B variants of the batch kernel (seg 5 object) as separate libraries under lib/ab/.
#   tools/batch_build.sh name "-DB3_ABL=1" [name2 "flags2" ...]   ->  lib/ab/libswn_b5_<name>.so (select with SWN_B200_LIB)
set -e
cd "$(dirname "$0")/../shallow-wavenet_b200/csrc"
OBJ=../lib/obj; AB=../lib/ab; mkdir -p $AB
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON=$(ls $OBJ/*.o | grep -v swn_batch_seg5.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DSWN_SEG=5 $flags -c swn_batch_inst.cu -o $AB/batch5_$name.o 2> $AB/batch5_$name.log \
    && nvcc $ARCH -shared -o $AB/libswn_b5_$name.so $COMMON $AB/batch5_$name.o -cudart static && echo "built $name" ) &
done
wait
