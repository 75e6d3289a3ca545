#!/bin/bash
mkdir -p gpurun_out/r02c
export SWN_PARITY_LOG=$PWD/gpurun_out/r02c/parity_forward.jsonl
rm -f $SWN_PARITY_LOG
timeout 600 python -m pytest tests/test_forward.py -m gpu -q -x > gpurun_out/r02c/pytest_forward.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02c/pytest_forward.log
cat $SWN_PARITY_LOG 2>/dev/null | head -20
