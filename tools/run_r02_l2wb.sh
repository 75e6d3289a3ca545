#!/bin/bash
# DRAM traffic of a 5.2 MB buffer rewritten in place 2000 times (10.6 GB of stores) by store / consumer-load flavour
cd /root/repo; mkdir -p gpurun_out/r02wb
for s in ${SMODES:-0 1 2 3}; do for l in ${LMODES:-0 1 2}; do
  timeout 120 ./tools/microbench/l2_writeback $s $l ${ROUNDS:-2000} ${SLEEP:-0} ${POLLUTE:-0} ${NTILES:-72} ${CYCLIC:-0} > gpurun_out/r02wb/plain_${s}_${l}.txt 2>&1 || { echo "plain run $s $l failed"; cat gpurun_out/r02wb/plain_${s}_${l}.txt; continue; }
  timeout 300 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02wb/ncu_${s}_${l}.csv ./tools/microbench/l2_writeback $s $l ${ROUNDS:-2000} ${SLEEP:-0} ${POLLUTE:-0} ${NTILES:-72} ${CYCLIC:-0} > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02wb/ncu_${s}_${l}.csv")) if len(r)>10 and r[0].isdigit()]
d={r[-3]:(float(r[-1].replace(",","")), r[-2]) for r in rows}
def gb(k):
    v,u=d[k]; return v*{"byte":1e-9,"Kbyte":1e-6,"Mbyte":1e-3,"Gbyte":1.0}.get(u,1e-9)
print("store %d load %d: dram write %8.3f GB  dram read %8.3f GB  L2 traffic %8.2f GB  | %s" % ($s,$l,gb("dram__bytes_write.sum"),gb("dram__bytes_read.sum"),gb("lts__t_bytes.sum"),open("gpurun_out/r02wb/plain_${s}_${l}.txt").read().strip()))
PY
done; done | tee -a gpurun_out/r02wb/summary.txt
