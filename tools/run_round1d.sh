# final bench lines of round 1 (after the v2 step-time work) + a source-level capture of the batch kernel
cd /root/repo
python bench.py --workload cfg3s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_cfg3s_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:swn_batch_kernel -c 1 -f -o gpurun_out/prof_v3d python bench.py --workload cfg3s --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_v3d.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01d.json 2> gpurun_out/bench_ref_r01d.err
python bench.py --workload cfg1 --steps 3 --warmup 3 > gpurun_out/bench_cfg1_r01d.json 2>/dev/null
python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_r01d.json 2>/dev/null
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg2_r01d.json 2>/dev/null
for f in gpurun_out/bench_*_r01d.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], d.get("impl"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
PY
done
