for v in 0 1 2 3; do
  export MPC_LIB_NAME=libmpc_bs$v.so
  for B in 16384 65536 1048576; do
    timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('bs=$v B=$B', r['kernel'], 'k_ms=%.3f value=%.4g mean_it=%.3f'%(r['kernel_ms'], d['value'], r['mean_iters']))"
  done
done
