"""Condense an .ncu-rep (ncu --set full) into the per-kernel table committed under profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.md
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math_pipe_throttle %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short_scoreboard %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name.split('(')[0]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in METRICS:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main()
