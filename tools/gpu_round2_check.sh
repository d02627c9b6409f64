set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/r2_gputest2.log
python bench.py --steps 10 --warmup 3 --cpu-seconds 8 > gpurun_out/r2_bench2.log 2>&1
python bench.py --steps 10 --warmup 3 --n-starts 1 --no-cpu-baseline --no-extras > gpurun_out/r2_bench2_s1.log 2>&1
python bench.py --config 1 > gpurun_out/r2_cfg1.log 2>&1
python bench.py --config 1 --n-starts 1 >> gpurun_out/r2_cfg1.log 2>&1
tail -5 gpurun_out/r2_gputest2.log
