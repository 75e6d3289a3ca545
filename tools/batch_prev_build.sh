#!/bin/bash
# Dev aid: the batch kernel (seg 5 object) of the last commit as lib/ab/libswn_b5_prev.so (A/B against the working tree)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
T=/tmp/oldcsrc_b; mkdir -p $T $ROOT/shallow-wavenet_b200/lib/ab
cp $ROOT/shallow-wavenet_b200/csrc/*.h $ROOT/shallow-wavenet_b200/csrc/*.cuh $ROOT/shallow-wavenet_b200/csrc/swn_batch_inst.cu $T/
git -C $ROOT show HEAD:shallow-wavenet_b200/csrc/swn_batch.cuh > $T/swn_batch.cuh
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DSWN_SEG=5 -c $T/swn_batch_inst.cu -o $ROOT/shallow-wavenet_b200/lib/ab/batch5_prev.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/shallow-wavenet_b200/lib/ab/libswn_b5_prev.so $(ls $ROOT/shallow-wavenet_b200/lib/obj/*.o | grep -v swn_batch_seg5.o) $ROOT/shallow-wavenet_b200/lib/ab/batch5_prev.o -cudart static
echo built prev
