#!/bin/bash
# Dev aid: the seg cluster kernel of the last commit as lib/ab/libswn_seg_prev.so (A/B against the working tree with tools/seg_ab_run.py)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
T=/tmp/oldcsrc; mkdir -p $T $ROOT/shallow-wavenet_b200/lib/ab
cp $ROOT/shallow-wavenet_b200/csrc/*.h $ROOT/shallow-wavenet_b200/csrc/*.cuh $ROOT/shallow-wavenet_b200/csrc/swn_coreseg_inst.cu $T/
for f in swn_coreseg.cuh swn_core.cuh swn_core2.cuh; do git -C $ROOT show HEAD:shallow-wavenet_b200/csrc/$f > $T/$f; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DSWN_NCH=3 -c $T/swn_coreseg_inst.cu -o $ROOT/shallow-wavenet_b200/lib/ab/coreseg3_prev.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/shallow-wavenet_b200/lib/ab/libswn_seg_prev.so $(ls $ROOT/shallow-wavenet_b200/lib/obj/*.o | grep -v swn_coreseg_nch3.o) $ROOT/shallow-wavenet_b200/lib/ab/coreseg3_prev.o -cudart static
echo built prev
