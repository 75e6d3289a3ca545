#!/bin/bash
# first GPU run of the seg > 1 cluster kernel: parity, then cfg1 timing against the layer-group kernel
mkdir -p gpurun_out/r02q
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seg_cluster" 2>&1 | tail -15 | tee gpurun_out/r02q/pytest_seg.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --workload cfg1 --steps 3 --warmup 3 > gpurun_out/r02q/bench_cfg1.json 2> gpurun_out/r02q/bench_cfg1.err
tail -c 1500 gpurun_out/r02q/bench_cfg1.json; tail -5 gpurun_out/r02q/bench_cfg1.err
