#!/bin/bash
mkdir -p gpurun_out/r02g
python tools/ab_run.py cfg2s > gpurun_out/r02g/ab.txt 2>&1; tail -5 gpurun_out/r02g/ab.txt
SWN_B200_LIB=/tmp/libswn_o15stamps.so python tools/gpu_prof_core.py cfg2s > gpurun_out/r02g/timeline_o15stamps.txt 2>&1; sed -n 2,24p gpurun_out/r02g/timeline_o15stamps.txt
for wl in cfg3s cfg5s; do
  for lib in default b3wt; do
    if [ $lib = b3wt ]; then export SWN_B200_LIB=$PWD/shallow-wavenet_b200/lib/ab/libswn_b3wt.so; else unset SWN_B200_LIB; fi
    python bench.py --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl $lib', round(d['value']), 'samples/s kernel_ms', round(d['kernel_ms'],3), 'variant', d['plan']['kernel_variant'])"
  done
done
unset SWN_B200_LIB
