"""Condense ncu reports (--set full) into a markdown table of the counters the design argues from.
usage: ncu_summary.py out.md rep1.ncu-rep [rep2 ...]"""
import csv, subprocess, sys, io, os
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short-scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg-throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
out = open(sys.argv[1], "w")
out.write("# ncu --set full --clock-control none captures (cold cache, one launch each)\n\n")
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out.write(f"## {os.path.basename(rep)}\n\n")
    kern = [r for r in rows[2:]]
    out.write("| counter | " + " | ".join(r[idx["Kernel Name"]].split("(")[0][:28] for r in kern) + " |\n")
    out.write("|---|" + "---|" * len(kern) + "\n")
    for k, label in KEYS:
        if k not in idx: continue
        out.write(f"| {label} ({units[idx[k]]}) | " + " | ".join(r[idx[k]] for r in kern) + " |\n")
    out.write("\n")
out.close()
print(open(sys.argv[1]).read())
