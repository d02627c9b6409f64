set -x
python -m pytest tests -m gpu -q -s -k "best_known or every_converged" 2>&1 | grep -E "n_starts|status 0|passed|failed" > gpurun_out/r2_gpu_parity_rates.log
for ns in 1 4; do
  python examples/train_a2c_mpc_batched.py --envs 1024 --updates 5 --n-steps 64 --horizon 16 --n-starts $ns >> gpurun_out/r2_config4.log 2>&1
done
python examples/train_a2c_mpc_batched.py --envs 1024 --updates 3 --n-steps 64 --horizon 16 --n-starts 1 --eager >> gpurun_out/r2_config4.log 2>&1
python examples/train_a2c_mpc_batched.py --envs 1024 --updates 3 --n-steps 64 --horizon 16 --n-starts 1 --algo ppo >> gpurun_out/r2_config4.log 2>&1
python examples/train_a2c_mpc_batched.py --envs 16384 --updates 3 --n-steps 64 --horizon 16 --n-starts 1 >> gpurun_out/r2_config4.log 2>&1
cat gpurun_out/r2_gpu_parity_rates.log
