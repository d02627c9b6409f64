set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ncu.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_solve_tmem -s 3 -c 1 -f -o gpurun_out/prof_solve_final $CMD > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 3 -c 1 -f -o gpurun_out/prof_prepare_final $CMD > gpurun_out/ncu_p.log 2>&1
ls -la gpurun_out/*.ncu-rep
