set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -x -q -m gpu > gpurun_out/r2a_pytest_scale.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_scale.log
python bench.py --steps 10 --warmup 3 --no-sweep --concurrent 0 > gpurun_out/r2a_bench_large.json 2> gpurun_out/r2a_bench_large.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r2a_ref_large.json 2> gpurun_out/r2a_ref_large.err
nproc > gpurun_out/r2a_nproc.txt; free -g >> gpurun_out/r2a_nproc.txt
tail -5 gpurun_out/r2a_pytest_scale.log
cat gpurun_out/r2a_bench_large.json | head -c 1500
