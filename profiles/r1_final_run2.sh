set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/fin2_small_sphere.json 2> gpurun_out/fin2_small_sphere.err
B="--no-cpu --concurrent 0"
python bench.py --workload large_sphere $B --sweep-n 512 > gpurun_out/fin2_large_sphere.json 2>/dev/null
python bench.py --workload medium_sphere $B --sweep-n 512 > gpurun_out/fin2_medium_sphere.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ --csv --log-file gpurun_out/fin2_launches_large.csv python bench.py --workload large_sphere --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
