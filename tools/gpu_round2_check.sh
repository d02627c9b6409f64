set -x
python -m pytest tests -m gpu -q -k "independent_of_batch or full_size or best_known" 2>&1 | tail -5 > gpurun_out/r2_gputest8.log
python bench.py --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2_bench8.log 2>&1
python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_bench8_s1.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_solve_tmem -s 3 -c 1 --csv --log-file gpurun_out/r2_ncu_bank.csv $CMD > gpurun_out/ncu_bank.log 2>&1
tail -3 gpurun_out/r2_gputest8.log; grep k_solve gpurun_out/r2_ncu_bank.csv | awk -F'","' '{print $(NF-2), $NF}'
