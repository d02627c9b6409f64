set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_gputest3.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.log 2>&1
python bench.py --steps 10 --warmup 3 --n-starts 1 --no-cpu-baseline --no-extras > gpurun_out/r2_bench3_s1.log 2>&1
tail -5 gpurun_out/r2_gputest3.log
