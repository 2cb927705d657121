"""one line per `ncu --page raw --csv` export: python scripts/ncu_raw_summary.py profiles/r02_final_ncu_*_raw.csv"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "winst"), ("smsp__issue_active.avg.pct", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "no_inst"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "no_inst/issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "short_sb/issue"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb/issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier/issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait/issue"),
        ("sm__icc_hit_rate.pct", "icc_hit%"), ("sm__inst_cache_hit_rate.pct", "icache_hit%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__bytes.sum.per_second", "dram/s"),
        ("lts__t_sectors_op_write.sum", "l2_wr_sectors")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h, units, v = rows[0], rows[1], rows[2]
    out = [path.split("/")[-1], v[h.index("Kernel Name")][:60] if "Kernel Name" in h else ""]
    for k, name in KEYS:
        if k in h:
            i = h.index(k)
            out.append(f"{name}={v[i]}{units[i] if name.startswith('dram') else ''}")
    print(" ".join(out))
