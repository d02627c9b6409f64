set -x
cp mpc-rl_for_avs_b200/libmpcb200.so /tmp/lib_main.so
CMD="python bench.py --steps 2 --warmup 3 --n-starts 1 --no-cpu-baseline --no-extras"
for v in main la0; do
  if [ $v = main ]; then cp /tmp/lib_main.so mpc-rl_for_avs_b200/libmpcb200.so; else cp mpc-rl_for_avs_b200/libmpcb200_$v.so mpc-rl_for_avs_b200/libmpcb200.so; fi
  python bench.py --steps 10 --n-starts 1 --no-extras --no-cpu-baseline > gpurun_out/r2_ab2_$v.log 2>&1
  $CMD > gpurun_out/plain_$v.log 2>&1 && ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio --clock-control none -k regex:k_solve_tmem -s 3 -c 1 --csv --log-file gpurun_out/r2_ncu_ab_$v.csv $CMD > gpurun_out/ncu_ab_$v.log 2>&1
done
cp /tmp/lib_main.so mpc-rl_for_avs_b200/libmpcb200.so
