#!/bin/bash
mkdir -p gpurun_out/r02f
python tools/ab_run.py cfg2s > gpurun_out/r02f/ab.txt 2>&1; tail -8 gpurun_out/r02f/ab.txt
SWN_B200_LIB=/tmp/libswn_o7stamps.so python tools/gpu_prof_core.py cfg2s > gpurun_out/r02f/timeline_o7stamps.txt 2>&1; sed -n 2,24p gpurun_out/r02f/timeline_o7stamps.txt
timeout 600 python -m pytest tests/test_forward.py -m gpu -x -q 2>&1 | tail -2
python tools/gpu_forward_bench.py cswnv 8 1000 3 > gpurun_out/r02f/forward_cswnv.json 2>gpurun_out/r02f/forward_cswnv.err; cut -c1-330 gpurun_out/r02f/forward_cswnv.json
python tools/gpu_forward_bench.py dswnv 8 1000 3 > gpurun_out/r02f/forward_dswnv.json 2>gpurun_out/r02f/forward_dswnv.err; cut -c1-330 gpurun_out/r02f/forward_dswnv.json
