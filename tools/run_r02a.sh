#!/bin/bash
# round 2, GPU call A: full GPU suite with the parity-error log, default bench (both arms, short), launch-mode probe
mkdir -p gpurun_out/r02a
export SWN_PARITY_LOG=$PWD/gpurun_out/r02a/parity_errors.jsonl
rm -f $SWN_PARITY_LOG
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/r02a/pytest.rc
tail -5 gpurun_out/r02a/pytest.log
unset SWN_PARITY_LOG
SWN_DEBUG_LAUNCH=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a/bench_cfg2.json 2> gpurun_out/r02a/bench_cfg2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a/bench_cfg2.json
grep -m3 "swn\]" gpurun_out/r02a/bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err; echo "ref rc=$?"
tail -c 900 gpurun_out/r02a/bench_ref.json
