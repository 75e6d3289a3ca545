#!/bin/bash
# full GPU suite + default bench + cfg1 bench + crossover after the seg cluster kernel work
mkdir -p gpurun_out/r02f1
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02f1/pytest_all.log
python bench.py > gpurun_out/r02f1/bench_default.json 2> gpurun_out/r02f1/bench_default.err; echo "bench rc=$?"
python bench.py --no-extras --no-cpu-baseline --workload cfg1 --steps 5 --warmup 3 > gpurun_out/r02f1/bench_cfg1.json 2> gpurun_out/r02f1/bench_cfg1.err
SWN_PROBE_FRAMES=60 python tools/gpu_crossover_probe.py 2>&1 | grep cswnv > gpurun_out/r02f1/crossover_cswnv.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f1/bench_default.json'))
print({k:d.get(k) for k in ('value','ms_per_step','rtf','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_each'], 'frac', d['roofline']['frac'])
print('batched', d['batched']['cfg3']['value'], d['batched']['cfg3']['e2e']['value'])
print('latency', d['latency']['cfg1']['value'], d['latency']['cfg1']['e2e']['value'], d['latency']['cfg1']['roofline']['frac'])
d=json.load(open('gpurun_out/r02f1/bench_cfg1.json')); print('cfg1', d['value'], d['e2e']['value'], d['roofline']['us_per_time_step'], d['roofline']['frac'])
PY
cat gpurun_out/r02f1/crossover_cswnv.txt | cut -c1-230 | head -5
