#!/bin/bash
# round 2 profile captures (each only after the same command has exited 0 without ncu)
mkdir -p gpurun_out/r02p
cd /root/repo
B="python bench.py --no-extras --no-cpu-baseline"
$B --workload cfg2s --steps 2 --warmup 3 > gpurun_out/r02p/plain_cfg2s.json 2> gpurun_out/r02p/plain_cfg2s.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p/launches_cfg2s.csv $B --workload cfg2s --steps 2 --warmup 3 > gpurun_out/r02p/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swn_core_kernel2 -s 3 -c 1 -f -o gpurun_out/r02p/prof_core2_cfg2s $B --workload cfg2s --steps 1 --warmup 3 > gpurun_out/r02p/ncu_core2.log 2>&1
# DRAM traffic of the default workload's kernel at full size
$B --steps 1 --warmup 3 > gpurun_out/r02p/plain_cfg2.json 2> gpurun_out/r02p/plain_cfg2.err && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:swn_core_kernel2 -s 3 -c 1 --csv --log-file gpurun_out/r02p/traffic_cfg2.csv $B --steps 1 --warmup 3 > gpurun_out/r02p/ncu_traffic.log 2>&1
# the forward GEMM (two CTAs per SM)
python tools/gpu_forward_bench.py cswnv 2 400 1 > gpurun_out/r02p/plain_fwd.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd_gemm -s 8 -c 2 -f -o gpurun_out/r02p/prof_fwd python tools/gpu_forward_bench.py cswnv 2 400 1 > gpurun_out/r02p/ncu_fwd.log 2>&1
ls -la gpurun_out/r02p
tail -3 gpurun_out/r02p/traffic_cfg2.csv
