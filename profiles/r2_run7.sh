# tnb_subpoly_batch: tests, the bench's concurrent leg (cluster of 16 / of 8 CTAs), then the whole GPU suite
set -x
mkdir -p gpurun_out
T=${TAG:-r4f}
python -m pytest tests/test_gpu_batch.py -x -q -m gpu > gpurun_out/${T}_pytest_batch.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_batch.log
python bench.py --workload small_sphere --steps 20 --warmup 3 --no-sweep --concurrent 8 --no-cpu > gpurun_out/${T}_bench_small_c16_k8.json 2> gpurun_out/${T}_bench_small_c16_k8.err
TNB_CLUSTER_CTAS=8 python bench.py --workload small_sphere --steps 20 --warmup 3 --no-sweep --concurrent 8 --no-cpu > gpurun_out/${T}_bench_small_c8_k8.json 2> gpurun_out/${T}_bench_small_c8_k8.err
TNB_CLUSTER_CTAS=8 python bench.py --workload small_sphere --steps 20 --warmup 3 --no-sweep --concurrent 16 --no-cpu > gpurun_out/${T}_bench_small_c8_k16.json 2> gpurun_out/${T}_bench_small_c8_k16.err
python bench.py --workload small_sphere --steps 20 --warmup 3 --no-sweep --concurrent 12 --no-cpu > gpurun_out/${T}_bench_small_c16_k12.json 2> gpurun_out/${T}_bench_small_c16_k12.err
python - <<PY
import json
for n in ("c16_k8","c8_k8","c8_k16","c16_k12"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_small_%s.json"%n)); c=d["concurrent"]; print(n, d["ms_per_step"], c["vs_one_at_a_time"], c["objects_per_s"])
    except Exception as e: print(n, "failed", e)
PY
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_all.log
tail -n 4 gpurun_out/${T}_pytest_batch.log gpurun_out/${T}_pytest_all.log
