mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-adaptive > gpurun_out/r03_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"restrict_filter|jump_fill|ce_kernel" -c 6 -o gpurun_out/r03_restrict -f python -m pytest tests/test_gpu_adapt.py -q -m gpu -k "test_leaf_coarsening_indicator_lifted and CDF44" > gpurun_out/r03_ncu_restrict.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_halo_launches_raw.csv python -m pytest tests/test_multi_halo.py -q -m gpu -k "test_rk4_on_a_graded_grid_across_ranks and 2-True" > gpurun_out/r03_ncu_halo.log 2>&1
ls -la gpurun_out | tail -8
