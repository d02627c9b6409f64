for v in 1 3 4; do
  if [ $v = 1 ]; then unset MPC_LIB_NAME; else export MPC_LIB_NAME=libmpc_pm$v.so; fi
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('min_blocks=$v', 'prep_ms=%.3f k_ms=%.3f value=%.4g'%(r['prepare_kernel_ms'], r['kernel_ms'], d['value']))"
  timeout 300 python -m pytest tests -m gpu -q -k "prepare or latch or collision or reference" 2>&1 | tail -1
done
