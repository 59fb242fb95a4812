"""Development tool: one line of key metrics per launch of an ncu report.  python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
want = [("gpu__time_duration.sum", "ms"), ("smsp__inst_executed.sum", "inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"), ("launch__occupancy_limit_registers", "occ_reg"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("launch__shared_mem_per_block_static", "smem_s"),
        ("launch__shared_mem_per_block_dynamic", "smem_d"), ("launch__grid_size", "grid"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st_barrier"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"), ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"), ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"), ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"), ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"), ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
        ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "st_sel"), ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "st_imc")]
ki = h.index("Kernel Name")
for r in rows[2:]:
    out = [r[ki].split("(")[0][-28:]]
    for w, n in want:
        if w in h:
            v = r[h.index(w)]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            out.append(f"{n}={v}")
    print(" ".join(out))
