# block cache A/B (TNB_NO_BLOCK_CACHE) on the small and the large sphere, K extractions in flight, parity
set -x
mkdir -p gpurun_out
T=${TAG:-r4e}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_repair.py -x -q -m gpu > gpurun_out/${T}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_parity.log
for NC in 1 0; do
  if [ $NC = 1 ]; then export TNB_NO_BLOCK_CACHE=1; else unset TNB_NO_BLOCK_CACHE; fi
  python bench.py --workload small_sphere --steps 20 --warmup 3 --no-sweep --concurrent 8 --no-cpu > gpurun_out/${T}_bench_small_nocache${NC}.json 2> gpurun_out/${T}_bench_small_nocache${NC}.err
  python bench.py --steps 10 --warmup 3 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_large_nocache${NC}.json 2> gpurun_out/${T}_bench_large_nocache${NC}.err
done
unset TNB_NO_BLOCK_CACHE
timeout 300 python tests/batch_phases.py small_sphere > gpurun_out/${T}_batch_phases.log 2>&1
python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "sphere or latched" > gpurun_out/${T}_pytest_scale.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_scale.log
tail -n 3 gpurun_out/${T}_pytest_parity.log gpurun_out/${T}_pytest_scale.log
python - <<PY
import json
for w in ("small","large"):
    for nc in (1,0):
        try:
            d=json.load(open("gpurun_out/${T}_bench_%s_nocache%d.json"%(w,nc))); print(w, "nocache" if nc else "cache  ", d["ms_per_step"], d["e2e"]["ms_per_step"], (d.get("concurrent") or {}).get("vs_one_at_a_time"), (d.get("concurrent") or {}).get("objects_per_s"))
        except Exception as e: print(w, nc, "failed", e)
PY
cat gpurun_out/${T}_batch_phases.log
