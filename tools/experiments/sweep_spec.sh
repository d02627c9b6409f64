for v in default 128_8 256_4 256_8; do
  if [ $v = default ]; then unset MPC_LIB_NAME; else export MPC_LIB_NAME=libmpc_sp_$v.so; fi
  for B in 65536 16384; do
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch $B 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('spec=$v B=$B', r['kernel'], 'k_ms=%.3f value=%.4g mean_it=%.6f'%(r['kernel_ms'], d['value'], r['mean_iters']))"
  done
done
