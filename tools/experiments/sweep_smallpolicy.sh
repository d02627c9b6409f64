for pol in 0 1; do
  export MPC_SMALL_POLICY=$pol
  for B in 256 1024 4096 16384 24576; do
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('policy=$pol B=$B', r['kernel'], 'k_ms=%.3f value=%.4g mean_it=%.6f'%(r['kernel_ms'], d['value'], r['mean_iters']))"
  done
done
unset MPC_SMALL_POLICY
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
