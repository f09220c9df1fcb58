# gradient-descent repair checks + curve-path regression + where K concurrent small extractions lose their overlap
set -x
mkdir -p gpurun_out
T=${TAG:-r4b}
python -m pytest tests/test_repair.py tests/test_abi.py -x -q -m gpu > gpurun_out/${T}_pytest_repair.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_repair.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/${T}_pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_parity.log
python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "torus" > gpurun_out/${T}_pytest_scale_torus.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_scale_torus.log
python bench.py --workload medium_torus --path curve --steps 10 --warmup 3 --no-sweep --concurrent 0 --no-cpu > gpurun_out/${T}_bench_medium_torus_curve.json 2> gpurun_out/${T}_bench_medium_torus_curve.err
timeout 300 python tests/batch_phases.py small_sphere > gpurun_out/${T}_batch_phases.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_repair.log gpurun_out/${T}_pytest_parity.log gpurun_out/${T}_pytest_scale_torus.log
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_medium_torus_curve.json")); print("medium torus curve", d["ms_per_step"], d["e2e"]["ms_per_step"])
PY
cat gpurun_out/${T}_batch_phases.log
