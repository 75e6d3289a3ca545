#!/bin/bash
mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02e/pytest.log
python tools/ab_run.py cfg2s > gpurun_out/r02e/ab.txt 2>&1; cat gpurun_out/r02e/ab.txt | tail -6
SWN_B200_LIB=/tmp/libswn_v2stamps.so python tools/gpu_prof_core.py cfg2s > gpurun_out/r02e/timeline_v2stamps.txt 2>&1; head -20 gpurun_out/r02e/timeline_v2stamps.txt
python tools/gpu_forward_bench.py cswnv 8 1000 3 > gpurun_out/r02e/forward_cswnv.json 2>gpurun_out/r02e/forward_cswnv.err; cut -c1-400 gpurun_out/r02e/forward_cswnv.json
