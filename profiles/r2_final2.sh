# the whole GPU suite and the default bench line after the last changes of the round
set -x
mkdir -p gpurun_out
T=${TAG:-r2fin2}
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_all.log
python bench.py > gpurun_out/${T}_bench_large_sphere.json 2> gpurun_out/${T}_bench_large_sphere.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
tail -n 4 gpurun_out/${T}_pytest_all.log gpurun_out/${T}_smoke.log
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_large_sphere.json")); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], json.dumps(d["roofline"]["fp32"]), json.dumps(d["concurrent"]["headline_workload"])[:300])
PY
