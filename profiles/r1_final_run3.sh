set -x
python bench.py > gpurun_out/fin3_small_sphere.json 2> gpurun_out/fin3_small_sphere.err
python bench.py --workload large_sphere --no-cpu --concurrent 0 --sweep-n 512 > gpurun_out/fin3_large_sphere.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ --csv --log-file gpurun_out/fin3_launches_large.csv python bench.py --workload large_sphere --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > /dev/null 2>&1
python bench.py --workload medium_torus --path curve --no-cpu --concurrent 0 --no-sweep > gpurun_out/fin3_medium_torus_curve.json 2>/dev/null
