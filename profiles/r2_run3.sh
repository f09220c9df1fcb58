# tests + large/small bench + launch list + (NCU=1) ncu --set full of the top kernels, exported to CSV on the box
set -x
mkdir -p gpurun_out
T=${TAG:-r2g}
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_parity.log
python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "${SCALE_K:-sphere or latched}" > gpurun_out/${T}_pytest_scale.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_scale.log
python bench.py --steps 10 --warmup 3 --sweep-n 512 --concurrent 0 --no-cpu > gpurun_out/${T}_bench_large.json 2> gpurun_out/${T}_bench_large.err
python bench.py --workload small_sphere --steps 10 --warmup 3 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_small.json 2> gpurun_out/${T}_bench_small.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_large.csv python profiles/one_extraction.py large_sphere planar 0 > gpurun_out/${T}_ncu.log 2>&1
if [ "${NCU:-0}" = "1" ]; then
  ncu --set full --import-source on --clock-control none --profile-from-start off \
      --kernel-name regex:"${NCU_K:-k_sweep_chunk|k_region_rows|k_pair_count_seg|new_vertices|k_sort_rows|k_sweep_signs}" \
      -o /tmp/${T}_full python profiles/one_extraction.py large_sphere planar 256 > gpurun_out/${T}_ncu_full.log 2>&1
  ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>/dev/null
  ncu -i /tmp/${T}_full.ncu-rep --page source --csv --kernel-name regex:"${NCU_SRC:-k_pair_count_seg|k_region_rows}" > gpurun_out/${T}_ncu_source.csv 2>/dev/null
  ls -la /tmp/${T}_full.ncu-rep gpurun_out/${T}_ncu_*.csv
fi
tail -n 3 gpurun_out/${T}_pytest_parity.log gpurun_out/${T}_pytest_scale.log
python - <<PY
import json
for w in ("large","small"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%w)); print(w, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["by_kernel_ms_per_step"])
        if d.get("eval_sweep"): print("sweep", d["eval_sweep"]["points_per_s"])
    except Exception as e: print(w, "failed", e)
PY
