"""Reads `ncu -i X.ncu-rep --page raw --csv` files and prints one CSV row per captured kernel launch with the counters
the round's notes cite.  usage: python scripts/ncu_summary.py raw1.csv [raw2.csv ...] > profiles/....csv"""
import csv
import sys

KEEP = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occ_limit_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "lsu_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__icc_request_hit_rate.pct", "icache_hit_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram_read_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sass__inst_executed_local_loads", "local_loads"),
    ("sass__inst_executed_local_stores", "local_stores"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
]
w = csv.writer(sys.stdout)
w.writerow(["source"] + [f"{short} [{{unit}}]" for _, short in KEEP])
first = True
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    idx = {name: head.index(name) for name, _ in KEEP if name in head}
    if first:
        sys.stdout.seek(0) if False else None
        w.writerow(["(units)"] + [units[idx[name]] if name in idx else "" for name, _ in KEEP])
        first = False
    for r in rows[2:]:
        w.writerow([path.split("/")[-1]] + [r[idx[name]] if name in idx else "" for name, _ in KEEP])
