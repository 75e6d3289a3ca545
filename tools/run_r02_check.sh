#!/bin/bash
# round-end style check: smoke, default bench line (with sub-records), reference arm
mkdir -p gpurun_out/r02chk
cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02chk/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02chk/smoke.log
( time python bench.py > gpurun_out/r02chk/bench_default.json 2> gpurun_out/r02chk/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02chk/bench_default.json'))
print({k:d.get(k) for k in ('value','ms_per_step','rtf','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('batched', {k:d['batched']['cfg3'][k] for k in ('value','kernel_ms')}, d['batched']['cfg3']['e2e']['value'])
print('latency', {k:d['latency']['cfg1'][k] for k in ('value','kernel_ms','rtf')}, d['latency']['cfg1']['e2e']['value'], d['latency']['cfg1']['roofline']['frac'])
print('cpu', d['cpu_baseline'])
PY
