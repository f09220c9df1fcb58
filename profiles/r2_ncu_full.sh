set -x
mkdir -p gpurun_out
T=${TAG:-r2f}
python profiles/one_extraction.py large_sphere planar 256 > gpurun_out/${T}_one.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none --profile-from-start off \
    --kernel-name regex:"k_sweep_chunk|k_region_rows|k_pair_count_seg|k_new_vertices|k_sort_rows|k_sweep_signs|k_pair_write_long_seg|k_vertex_outputs|k_scan_count_mask|k_fan_scan" \
    -o gpurun_out/${T}_full python profiles/one_extraction.py large_sphere planar 256 > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_full.ncu-rep
tail -n 3 gpurun_out/${T}_ncu.log
