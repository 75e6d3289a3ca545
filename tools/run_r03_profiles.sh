#!/bin/bash
# round 2 (late) profile captures of the seg cluster kernel with three utterances side by side
# (each only after the same command has exited 0 without ncu)
mkdir -p gpurun_out/r03p
cd /root/repo
B="python bench.py --no-extras --no-cpu-baseline"
$B --workload cfg1x3 --steps 2 --warmup 3 > gpurun_out/r03p/plain_cfg1x3.json 2> gpurun_out/r03p/plain_cfg1x3.err || exit 1
$B --workload cfg1x7 --steps 3 --warmup 3 > gpurun_out/r03p/plain_cfg1x7.json 2> gpurun_out/r03p/plain_cfg1x7.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03p/launches_cfg1x3.csv $B --workload cfg1x3 --steps 2 --warmup 3 > gpurun_out/r03p/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swn_coreseg_kernel -s 3 -c 1 -f -o gpurun_out/r03p/prof_coreseg_cfg1x3 $B --workload cfg1x3 --steps 1 --warmup 3 > gpurun_out/r03p/ncu_coreseg.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:swn_coreseg_kernel -s 3 -c 1 --csv --log-file gpurun_out/r03p/traffic_cfg1x3.csv $B --workload cfg1x3 --steps 1 --warmup 3 > gpurun_out/r03p/ncu_traffic.log 2>&1
ls -la gpurun_out/r03p
tail -3 gpurun_out/r03p/traffic_cfg1x3.csv
tail -c 700 gpurun_out/r03p/plain_cfg1x3.json
