# DRAM traffic of the default workload's kernel (one --set full capture, after a plain run exited 0)
cd /root/repo
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_cfg2_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:swn_core_kernel -c 1 -f -o gpurun_out/prof_v2e_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v2e.log 2>&1
python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_cfg3_e.log 2>&1 && \
ncu --set full --clock-control none -k regex:swn_batch_kernel -c 1 -f -o gpurun_out/prof_v3e_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v3e.log 2>&1
ncu -i gpurun_out/prof_v2e_cfg2.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/traffic_cfg2.csv
ncu -i gpurun_out/prof_v3e_cfg3.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/traffic_cfg3.csv
cat gpurun_out/traffic_cfg2.csv gpurun_out/traffic_cfg3.csv
