set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2_r01c.json 2> gpurun_out/bench_cfg2_r01c.err; tail -c 600 gpurun_out/bench_cfg2_r01c.err
python bench.py --workload cfg2s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_cfg2s_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2c.csv python bench.py --workload cfg2s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_v2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swn_core_kernel -c 1 -f -o gpurun_out/prof_v2c python bench.py --workload cfg2s --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_v2c.log 2>&1
python tools/gpu_prof_core.py cfg2s 2>&1 | grep -v "variant=1\|^      \|layer-3\|head of warp" > gpurun_out/timeline_v2c.txt
ls -la gpurun_out | tail -8
