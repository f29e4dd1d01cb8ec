# round 2 capture commands (run on the GPU box through gpurun; outputs in gpurun_out/, summaries copied to profiles/r5_*)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r5_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-adaptive > gpurun_out/r5_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stage_kernel" -s 12 -c 4 -o gpurun_out/r5_stage -f python bench.py --steps 2 --warmup 3 --no-cpu --no-adaptive --no-wavelet --level 4 --e2e-trees 1 > gpurun_out/r5_ncu_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wavelet_fast_kernel" -s 3 -c 1 -o gpurun_out/r5_wavelet -f python bench.py --steps 1 --warmup 3 --no-cpu --no-adaptive --level 4 --e2e-trees 1 > gpurun_out/r5_ncu_wavelet.log 2>&1
ls -la gpurun_out | tail -n 8
