#!/bin/bash
# round 2, GPU call B: decode driver tests, bench with sub-records (N=1 incl. the sharded N=1 figure)
mkdir -p gpurun_out/r02b
timeout 600 python -m pytest tests/test_decode_driver.py -m gpu -x -q > gpurun_out/r02b/pytest_decode.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02b/pytest_decode.log
timeout 1200 python bench.py --steps 3 --warmup 3 --sharded-at-n1 --no-cpu-baseline > gpurun_out/r02b/bench_n1_sharded.json 2> gpurun_out/r02b/bench_n1_sharded.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02b/bench_n1_sharded.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b/bench_n1_sharded.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['e2e'])
b=d['batched']['cfg3']; print('cfg3', b['value'], b['e2e']['value'], b['kernel_ms'], b['roofline']['frac'], b['roofline']['frac_of_tf32x3_equivalent'])
s=d['sharded']['cfg4']; print('cfg4', {k:v for k,v in s.items() if k not in ('rank0','workload')})
PY
