#!/bin/bash
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/ab_run.py cfg2s > gpurun_out/r02m/ab.txt 2>&1; tail -3 gpurun_out/r02m/ab.txt
python tools/gpu_core_columns_probe.py 2>&1 | tee gpurun_out/r02m/columns.txt | grep "cluster kernel"
