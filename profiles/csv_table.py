"""Condense an `ncu --page raw --csv` dump into one row per kernel.   python profiles/csv_table.py in.csv [out.md]"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_lsu.sum", "lsu"), ("smsp__inst_executed.sum", "winst"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ldsec"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ldreq"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_ssb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("| kernel | " + " | ".join(f"{n} ({units[idx[m]]})" if m in idx else n for m, n in COLS) + " |\n")
out.write("|---|" + "---|" * len(COLS) + "\n")
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.replace("void ", "").split("(")[0][:48]
    vals = []
    for m, n in COLS:
        v = r[idx[m]] if m in idx else "-"
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.4g}"
        except ValueError:
            pass
        vals.append(v)
    out.write(f"| `{name}` | " + " | ".join(vals) + " |\n")
