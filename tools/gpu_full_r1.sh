set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench15.json 2> gpurun_out/bench15_err.log; tail -c 600 gpurun_out/bench15_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pipe_row_kernel|norm3_kernel|combine_kernel" -s 12 -c 4 -f -o gpurun_out/prof_r1_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-copy-floor > gpurun_out/ncu_full_run.log 2>&1
timeout 900 python tools/microbench.py --out gpurun_out/sweep_r1.jsonl > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
