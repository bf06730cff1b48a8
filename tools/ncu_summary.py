"""Prints the key metrics of every kernel in an .ncu-rep (raw page) -- used to write profiles/*.txt."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "not_selected", "barrier", "dispatch_stall", "branch_resolving", "no_instruction", "tex_throttle", "membar"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("== %s" % r[col["Kernel Name"]][:110])
        for key, name in WANT:
            if key in col:
                print("   %-18s %s %s" % (name, r[col[key]], units[col[key]]))
        st = []
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if key in col:
                st.append((float(r[col[key]]), s))
        print("   stalls/issue       " + "  ".join("%s=%.2f" % (s, v) for v, s in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
