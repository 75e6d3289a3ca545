#!/bin/bash
mkdir -p gpurun_out/r02h
python tools/ab_run.py cfg2s > gpurun_out/r02h/ab.txt 2>&1; tail -3 gpurun_out/r02h/ab.txt
SWN_B200_LIB=/tmp/libswn_stamps.so python tools/gpu_prof_core.py cfg2s > gpurun_out/r02h/timeline_stamps.txt 2>&1; sed -n 2,30p gpurun_out/r02h/timeline_stamps.txt
