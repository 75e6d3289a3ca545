#!/bin/bash
mkdir -p gpurun_out/r02y
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seg_cluster" 2>&1 | tail -15 | tee gpurun_out/r02y/pytest_seg.log
SWN_B200_LIB=/root/repo/shallow-wavenet_b200/lib/ab/libswn_seg_stamps.so python tools/gpu_prof_coreseg.py cfg1 > gpurun_out/r02y/timeline.txt 2>&1
cat gpurun_out/r02y/timeline.txt | tail -30
