# ncu launch list + one --set full capture of the four hot kernels (no tests, no sweep)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pipe_row_kernel|norm3_kernel|combine_kernel" -s 12 -c 4 -f -o gpurun_out/prof_r1_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out/prof_r1_final.ncu-rep gpurun_out/launches_r1.csv
