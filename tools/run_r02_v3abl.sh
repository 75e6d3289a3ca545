#!/bin/bash
# which global writes of the batch kernel reach DRAM?  (cfg3s, ncu dram metrics; ablated builds give wrong samples)
cd /root/repo; mkdir -p gpurun_out/r02abl
B="python bench.py --no-extras --no-cpu-baseline --workload cfg3s --steps 1 --warmup 1"
for v in default abl1 abl2; do
  if [ $v != default ]; then export SWN_B200_LIB=/root/repo/shallow-wavenet_b200/lib/ab/libswn_b5_$v.so; else unset SWN_B200_LIB; fi
  timeout 300 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:swn_batch_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02abl/ncu_$v.csv $B > gpurun_out/r02abl/run_$v.log 2>&1
  echo "== $v"; grep "swn_batch_kernel" gpurun_out/r02abl/ncu_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done | tee gpurun_out/r02abl/summary.txt
