set -x
mkdir -p gpurun_out
T=${TAG:-r2b}
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_parity.log
python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "sphere or planar or latched" > gpurun_out/${T}_pytest_scale.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_scale.log
python bench.py --steps 10 --warmup 3 --sweep-n 512 --concurrent 0 --no-cpu > gpurun_out/${T}_bench_large.json 2> gpurun_out/${T}_bench_large.err
python bench.py --workload small_sphere --steps 10 --warmup 3 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_small.json 2> gpurun_out/${T}_bench_small.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ --csv --log-file gpurun_out/${T}_launches_large.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > gpurun_out/${T}_ncu.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_parity.log gpurun_out/${T}_pytest_scale.log
python - <<PY
import json
for w in ("large","small"):
    if w=="large": print("sweep", json.load(open("gpurun_out/${T}_bench_large.json"))["eval_sweep"]["points_per_s"])
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%w)); print(w, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["by_kernel_ms_per_step"])
    except Exception as e: print(w, "failed", e)
PY
