"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small markdown table.

    python profiles/summarize_ncu.py gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    groups = {}
    for r in rows[2:]:
        groups.setdefault(r[idx["Kernel Name"]], []).append(r)
    with open(out, "w") as f:
        f.write(f"# ncu summary of `{rep}` (ncu --set full --clock-control none; cold-cache, serialised launches)\n\n")
        for name, rs in groups.items():
            f.write(f"## `{name[:110]}`  ({len(rs)} launches captured; first shown, min/max duration over all)\n\n")
            d = [float(r[idx['gpu__time_duration.sum']]) for r in rs]
            f.write(f"duration min/max: {min(d):.2f} / {max(d):.2f} {units[idx['gpu__time_duration.sum']]}; "
                    f"grid {rs[0][idx['Grid Size']]} x block {rs[0][idx['Block Size']]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m, label in METRICS:
                if m in idx:
                    f.write(f"| {label} (`{m}`) | {rs[0][idx[m]]} | {units[idx[m]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
