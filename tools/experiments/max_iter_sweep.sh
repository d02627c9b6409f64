for mi in 60 90 120; do
for S in 4 1; do
MPC_MAX_ITER=$mi python bench.py --steps 10 --n-starts $S --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('max_iter $mi S $S ms', round(j['ms_per_step'],3), 'conv', round(j['solver']['converged_frac'],4), 'settled', round(j['solver']['settled_frac'],4), 'cap', round(j['solver']['max_iter_frac'],4), 'mean_it', round(j['solver']['mean_iters'],2))"
done; done
