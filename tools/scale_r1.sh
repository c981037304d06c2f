# bench.py at N = 4 and N = 8 on one 8-GPU box (N = 1, 2 are run on smaller boxes); outputs under gpurun_out/
for N in 4 8; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n${N}_err.log
  tail -c 200 gpurun_out/bench_n${N}_err.log
done
timeout 300 python -m pytest tests/test_distributed_gpu.py -q -m gpu 2>&1 | tail -3
