"""Summarise an .ncu-rep (read here, without a GPU): per-launch duration, DRAM/L2 traffic, utilisation and
warp-stall breakdown.   usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--source KERNEL_SUBSTR]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.per_cycle_active", "issue/cycle"),
    ("smsp__inst_executed.sum", "instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "wait", "math_pipe_throttle", "lg_throttle",
          "membar", "sleeping", "dispatch_stall", "branch_resolving", "no_instruction", "not_selected", "tex_throttle", "drain", "imc_miss"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    ki = hdr.index("Kernel Name")
    for r in rows:
        name = r[ki].replace("<unnamed>::", "").replace("void ", "")
        print("== " + name[:90])
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("   %-22s %s %s" % (label, r[i], units[i]))
        st = []
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if key in hdr:
                v = float(r[hdr.index(key)] or 0)
                if v > 0.05:
                    st.append((v, s))
        print("   stalls/issue: " + ", ".join("%s %.2f" % (s, v) for v, s in sorted(st, reverse=True)))


if __name__ == "__main__":
    main()
