set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for B in 1024 4096 16384 24576 32768; do
  for T in 0 256; do
    if [ $T = 0 ]; then unset MPC_TPB; else export MPC_TPB=$T; fi
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('B=$B TPB=$T', r['kernel'], 'k_ms=%.3f prep=%.3f value=%.3g e2e=%.3g'%(r['kernel_ms'], r['prepare_kernel_ms'], d['value'], d['e2e']['value']))"
  done
done
