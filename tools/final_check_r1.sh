# what the driver does at round end, on one fresh box: GPU tests, smoke(), reference arm, own arm
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref_err.log; tail -c 400 gpurun_out/final_ref.json
timeout 600 python bench.py > gpurun_out/final_own.json 2> gpurun_out/final_own_err.log; tail -c 300 gpurun_out/final_own_err.log; head -c 600 gpurun_out/final_own.json
