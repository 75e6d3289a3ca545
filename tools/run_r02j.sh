#!/bin/bash
mkdir -p gpurun_out/r02j
python tools/ab_run.py cfg2s > gpurun_out/r02j/ab.txt 2>&1; tail -4 gpurun_out/r02j/ab.txt
SWN_B200_LIB=/tmp/libswn_recstamps.so python tools/gpu_prof_core.py cfg2s > gpurun_out/r02j/timeline_stamps.txt 2>&1; sed -n 2,28p gpurun_out/r02j/timeline_stamps.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or variant or fixture or philox or small_batches or zero_noise or recipe_size" 2>&1 | tail -2
