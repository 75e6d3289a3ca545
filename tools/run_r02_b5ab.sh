#!/bin/bash
# A/B of the batch kernel on cfg3s / cfg3 (working tree vs lib/ab/libswn_b5_prev.so) + parity of the batch tests
cd /root/repo; mkdir -p gpurun_out/r02b5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch or v3 or recipe_size_matches or large" 2>&1 | tail -2
for rep in 1 2; do
for v in default prev; do
  if [ $v != default ]; then export SWN_B200_LIB=/root/repo/shallow-wavenet_b200/lib/ab/libswn_b5_$v.so; else unset SWN_B200_LIB; fi
  python bench.py --no-extras --no-cpu-baseline --workload ${WL:-cfg3s} --steps 3 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$v', d['value'], d['kernel_ms'], d['roofline'].get('us_per_loop_iteration'))"
done; done | tee gpurun_out/r02b5/ab_${WL:-cfg3s}.txt
