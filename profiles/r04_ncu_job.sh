mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r04_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-adaptive > gpurun_out/r04_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stage_kernel" -c 4 -o gpurun_out/r04_stage_sphere -f python profiles/r04_penalized.py > gpurun_out/r04_ncu_sphere.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r04_cyl2d_launches_raw.csv python -m pytest tests/test_gpu_cylinder2d.py -q -m gpu -k "test_cylinder_fixture_2d and CDF44 and not norm and not significant" > gpurun_out/r04_ncu_cyl2d.log 2>&1
ls -la gpurun_out | tail -n 8
