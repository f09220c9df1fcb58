# closing evidence of round 2: bench lines (default run = large sphere with every leg), the other workloads, launch list,
# ncu --set full of the top kernels (metrics exported to CSV on the box)
set -x
mkdir -p gpurun_out
T=${TAG:-r2fin}
python bench.py > gpurun_out/${T}_bench_large_sphere.json 2> gpurun_out/${T}_bench_large_sphere.err
python bench.py --workload small_sphere --no-sweep --no-cpu > gpurun_out/${T}_bench_small_sphere.json 2> gpurun_out/${T}_bench_small_sphere.err
python bench.py --workload medium_torus --path curve --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_medium_torus_curve.json 2> gpurun_out/${T}_bench_medium_torus_curve.err
python bench.py --workload large_torus --steps 10 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_large_torus.json 2> gpurun_out/${T}_bench_large_torus.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_large.csv python profiles/one_extraction.py large_sphere planar 0 > gpurun_out/${T}_ncu.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off \
    --kernel-name regex:"k_sweep_pieces|k_pair_count_seg|k_sd_new_vertices|k_region_rows|k_sd_pair_long|k_sd_keep_write" \
    -o /tmp/${T}_full python profiles/one_extraction.py large_sphere planar 0 > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>/dev/null
ls -la /tmp/${T}_full.ncu-rep gpurun_out/${T}_ncu_*.csv
python - <<PY
import json
for w in ("large_sphere","small_sphere","medium_torus_curve","large_torus"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%w)); print(w, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], json.dumps(d.get("concurrent"))[:600])
    except Exception as e: print(w, "failed", e)
PY
