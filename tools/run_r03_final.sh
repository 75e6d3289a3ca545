#!/bin/bash
# round 2 (late) final evidence: default bench, ncu launch list + full capture of the final one-sample-per-step cluster kernel (cfg2s)
mkdir -p gpurun_out/r03final
cd /root/repo
python bench.py > gpurun_out/r03final/bench_default.json 2> gpurun_out/r03final/bench_default.err || exit 1
B="python bench.py --no-extras --no-cpu-baseline"
$B --workload cfg2s --steps 2 --warmup 3 > gpurun_out/r03final/plain_cfg2s.json 2> gpurun_out/r03final/plain_cfg2s.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03final/launches_cfg2s.csv $B --workload cfg2s --steps 2 --warmup 3 > gpurun_out/r03final/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swn_core_kernel2 -s 3 -c 1 -f -o gpurun_out/r03final/prof_core2_cfg2s $B --workload cfg2s --steps 1 --warmup 3 > gpurun_out/r03final/ncu_core2.log 2>&1
tail -c 300 gpurun_out/r03final/bench_default.json
ls -la gpurun_out/r03final
