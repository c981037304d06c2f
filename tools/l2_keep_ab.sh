python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "norm3 or combine or k4 or clip or grad" 2>&1 | tail -3
for mb in 0 80 110 48; do
  SISS_L2_KEEP_MB=$mb python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-extra-configs --no-copy-floor 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['roofline']['kernels']
print('keep_mb', $mb, 'ms/step', round(d['ms_per_step'], 4), 'norm3', round(k['siss_norm3']['ms'] * 1e3, 1), 'combine', round(k['siss_combine']['ms'] * 1e3, 1), 'value', round(d['value']))
"
done
