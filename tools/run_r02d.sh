#!/bin/bash
mkdir -p gpurun_out/r02d
python tools/gpu_forward_bench.py cswnv 8 1000 5 > gpurun_out/r02d/forward_cswnv.json 2> gpurun_out/r02d/forward_cswnv.err; echo "fwd cswnv rc=$?"; cat gpurun_out/r02d/forward_cswnv.json
python tools/gpu_forward_bench.py dswnv 8 1000 5 > gpurun_out/r02d/forward_dswnv.json 2> gpurun_out/r02d/forward_dswnv.err; echo "fwd dswnv rc=$?"; cat gpurun_out/r02d/forward_dswnv.json
ncu --set full --clock-control none --import-source on -k regex:fwd_gemm -s 8 -c 2 -o gpurun_out/r02d/prof_fwd python tools/gpu_forward_bench.py cswnv 2 400 1 > gpurun_out/r02d/ncu_fwd.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r02d/bench_cfg5_full.json 2> gpurun_out/r02d/bench_cfg5_full.err; echo "cfg5 rc=$?"
tail -c 1200 gpurun_out/r02d/bench_cfg5_full.json; tail -3 gpurun_out/r02d/bench_cfg5_full.err
