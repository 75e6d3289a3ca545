#!/bin/bash
# round-end evidence: full GPU suite, smoke, default bench, launch list + ncu capture of the seg cluster kernel (cfg1) and of the cluster kernel (cfg2s)
mkdir -p gpurun_out/r02f2
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02f2/pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f2/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r02f2/bench_default.json 2> gpurun_out/r02f2/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f2/bench_reference.json 2> gpurun_out/r02f2/bench_reference.err; echo "ref rc=$?"
B="python bench.py --no-extras --no-cpu-baseline"
$B --workload cfg1 --steps 2 --warmup 3 > gpurun_out/r02f2/plain_cfg1.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f2/launches_cfg1.csv $B --workload cfg1 --steps 2 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:swn_coreseg_kernel -s 3 -c 1 -f -o gpurun_out/r02f2/prof_coreseg_cfg1 $B --workload cfg1 --steps 1 --warmup 3 > gpurun_out/r02f2/ncu_coreseg.log 2>&1
$B --workload cfg2s --steps 2 --warmup 3 > gpurun_out/r02f2/plain_cfg2s.json 2> /dev/null && \
ncu --set full --clock-control none --import-source on -k regex:swn_core_kernel2 -s 3 -c 1 -f -o gpurun_out/r02f2/prof_core2_cfg2s $B --workload cfg2s --steps 1 --warmup 3 > gpurun_out/r02f2/ncu_core2.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:swn_coreseg_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02f2/traffic_cfg1.csv $B --workload cfg1 --steps 1 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f2/bench_default.json'))
print({k:d.get(k) for k in ('value','ms_per_step','rtf','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('batched', d['batched']['cfg3']['value'], 'latency', d['latency']['cfg1']['value'])
r=json.load(open('gpurun_out/r02f2/bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
tail -3 gpurun_out/r02f2/traffic_cfg1.csv | awk -F'","' '{print $(NF-2), $NF}'
ls gpurun_out/r02f2
