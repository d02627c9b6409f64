set -x
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_gputest7.log
python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2_bench7.log 2>&1
python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench7_s1.log 2>&1
tail -3 gpurun_out/r2_gputest7.log
