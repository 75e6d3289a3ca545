#!/bin/bash
mkdir -p gpurun_out/r02t
cd /root/repo
SWN_B200_LIB=/root/repo/shallow-wavenet_b200/lib/ab/libswn_seg_stamps.so python tools/gpu_prof_coreseg.py cfg1 > gpurun_out/r02t/timeline.txt 2>&1
cat gpurun_out/r02t/timeline.txt | tail -30
