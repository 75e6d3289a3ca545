#!/bin/bash
mkdir -p gpurun_out/r02y2
cd /root/repo
B="python bench.py --no-extras --no-cpu-baseline --workload cfg1"
$B --steps 2 --warmup 3 > gpurun_out/r02y2/plain_cfg1.json 2> gpurun_out/r02y2/plain_cfg1.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:swn_coreseg_kernel -s 3 -c 1 -f -o gpurun_out/r02y2/prof_coreseg_cfg1 $B --steps 1 --warmup 3 > gpurun_out/r02y2/ncu.log 2>&1
tail -2 gpurun_out/r02y2/ncu.log
