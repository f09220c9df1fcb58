mkdir -p gpurun_out
for cfg in "2 4" "4 4" "6 4" "4 8"; do
  set -- $cfg
  TNB_PAIR_LONG_CTAS=$1 TNB_WIDE_ROW_CTAS=$2 python bench.py --steps 20 --warmup 3 --no-sweep --concurrent 0 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', d['ms_per_step'], d['roofline']['by_kernel_ms_per_step']['step_back'], d['roofline']['by_kernel_ms_per_step']['face_rows'])"
done
