set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_gputest_final.log
python bench.py > gpurun_out/r2_bench_final.log 2>&1
python bench.py --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_final_s1.log 2>&1
python bench.py --batch 1048576 --steps 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_1m.log 2>&1
python bench.py --n-starts 1 --batch 1048576 --steps 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_1m_s1.log 2>&1
python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r2_bench_reference.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_ncu.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve_tmem -s 3 -c 1 -o gpurun_out/r02_prof_solve -f $CMD > gpurun_out/ncu_solve.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 3 -c 1 -o gpurun_out/r02_prof_prepare -f $CMD > gpurun_out/ncu_prepare.log 2>&1
tail -5 gpurun_out/r2_gputest_final.log
