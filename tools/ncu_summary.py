"""Turns `ncu -i X.ncu-rep --page raw --csv` into the small JSON summaries kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_summary.py /tmp/raw.csv out.json [launch_index]
"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__thread_inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    data = rows[hdr + 2 + which]
    out = {"kernel": data[names.index("Kernel Name")] if "Kernel Name" in names else ""}
    for k in KEEP:
        if k in names:
            i = names.index(k)
            out[k] = {"unit": units[i], "value": data[i]}

    def val(k):
        v = out.get(k, {}).get("value")
        return float(v.replace(",", "")) if v else None
    r, w = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    if r is not None and w is not None:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        r *= scale.get(out["dram__bytes_read.sum"]["unit"], 1)
        w *= scale.get(out["dram__bytes_write.sum"]["unit"], 1)
        out["dram_traffic_bytes_per_launch"] = r + w
    bl, wl = val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"), val("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum")
    if bl is not None and wl:
        out["shared_ld_bank_conflict_fraction"] = bl / wl
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("kernel", "dram_traffic_bytes_per_launch", "shared_ld_bank_conflict_fraction") if k in out}))


if __name__ == "__main__":
    main()
