set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gputest9.log
python -m pytest tests -m gpu -q -s -k "best_known or every_converged" 2>&1 | grep -E "n_starts|status 0|passed|failed" > gpurun_out/r2_gpu_parity_rates.log
python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2_bench9.log 2>&1
python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench9_s1.log 2>&1
tail -8 gpurun_out/r2_gputest9.log; cat gpurun_out/r2_gpu_parity_rates.log; python - <<'PY'
import json
for f in ("gpurun_out/r2_bench9.log","gpurun_out/r2_bench9_s1.log"):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["ms_per_step"], j["value"], j["e2e"]["value"], j["solver"]["converged_frac"], j["solver"]["mean_iters"])
PY
