set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/fin_small_sphere.json 2> gpurun_out/fin_small_sphere.err
B="--no-cpu --concurrent 0"
python bench.py --workload medium_sphere $B --sweep-n 512 > gpurun_out/fin_medium_sphere.json 2>/dev/null
python bench.py --workload large_sphere $B --sweep-n 512 > gpurun_out/fin_large_sphere.json 2>/dev/null
python bench.py --workload medium_torus --path curve $B --no-sweep > gpurun_out/fin_medium_torus_curve.json 2>/dev/null
python bench.py --workload small_torus --path curve $B --no-sweep > gpurun_out/fin_small_torus_curve.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ --csv --log-file gpurun_out/fin_launches_small.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_ --csv --log-file gpurun_out/fin_launches_curve.csv python bench.py --workload medium_torus --path curve --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_steps_grid_curve -s 3 -c 1 -o gpurun_out/fin_steps_curve -f python bench.py --workload medium_torus --path curve --steps 1 --warmup 3 --no-cpu --no-sweep --concurrent 0 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_grid_train -s 30 -c 3 -o gpurun_out/fin_grid_train -f python tests/train_throughput.py small > /dev/null 2>&1
python tests/train_throughput.py small 2>&1 | grep iteration | tee gpurun_out/train_throughput.log
python tests/scale_check.py large_sphere 2>&1 | tail -3 | tee gpurun_out/scale_check_large.log
ls -la gpurun_out/*.ncu-rep
