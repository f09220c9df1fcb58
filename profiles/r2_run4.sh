# A/B of the sweep kernel's per-trip barrier + the usual validation
set -x
mkdir -p gpurun_out
T=${TAG:-r2o}
for SYNC in 0 1; do
  TNB_SWEEP_SYNC=$SYNC python bench.py --steps 10 --warmup 3 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_large_sync${SYNC}.json 2> gpurun_out/${T}_bench_large_sync${SYNC}.err
done
python - <<PY
import json
for s in (0,1):
    d=json.load(open("gpurun_out/${T}_bench_large_sync%d.json"%s)); print("sync",s, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["by_kernel_ms_per_step"])
PY
TAG=$T bash profiles/r2_run3.sh
