set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_gputest5.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench5.log 2>&1
python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench5_s1.log 2>&1
cp mpc-rl_for_avs_b200/libmpcb200.so /tmp/lib_main.so
cp mpc-rl_for_avs_b200/libmpcb200_pb2.so mpc-rl_for_avs_b200/libmpcb200.so
python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench5_s1_pb2.log 2>&1
cp /tmp/lib_main.so mpc-rl_for_avs_b200/libmpcb200.so
tail -5 gpurun_out/r2_gputest5.log
