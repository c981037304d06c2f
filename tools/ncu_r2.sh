#!/bin/bash
# round-2 ncu captures (1 GPU): launch list of the bench step, --set full of the four hot kernels, --set full of the rest
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pipe_row_kernel|norm3_kernel|combine_kernel" -s 12 -c 4 -f -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_full_run.log 2>&1
python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1; grep ALG_BYTES gpurun_out/ncu_targets_plain.log | sed 's/^ALG_BYTES //' > gpurun_out/r2_ncu_alg_bytes.json
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"^(pipe_row_kernel|add_noise_kernel|mixture_kernel|dual_mse|wmse_fwd_kernel|batch_stats|randn_kernel|draw_rows|combine_adamw|mt_|membership|sqerr|counter)" -f -o gpurun_out/prof_r2_rest python tools/ncu_targets.py > gpurun_out/ncu_rest_run.log 2>&1
ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/prof_r2_rest.ncu-rep gpurun_out/launches_r2.csv; tail -3 gpurun_out/ncu_rest_run.log
# keep what travels back small (gpurun merges at most 64 MiB): export the raw pages here, drop the big report
ncu -i gpurun_out/prof_r2_final.ncu-rep --page raw --csv > gpurun_out/prof_r2_final_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2_rest.ncu-rep --page raw --csv > gpurun_out/prof_r2_rest_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r2_rest.ncu-rep
ls -la gpurun_out/*.csv
