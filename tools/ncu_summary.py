#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [--src KERNEL_REGEX] -- text summary of an ncu report (metrics the roofline cites + stall mix)."""
import csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]

def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        print("=== %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, r[i], units[i]))
        st = [(hdr[i], r[i]) for i in range(len(hdr)) if "issue_stalled" in hdr[i] and "per_issue_active" in hdr[i] and "not_issued" not in hdr[i]]
        st = sorted(st, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:8]
        print("  stall cycles per issued instruction: " + ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(v)) for k, v in st))
    if "--src" in sys.argv:
        pat = sys.argv[sys.argv.index("--src") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        isrc, isamp, iexec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        sc = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        data = [r for r in rows[2:] if len(r) == len(h) and (r[isamp] or '0').isdigit()]
        tot = sum(int(r[isamp] or 0) for r in data)
        print("--- hottest SASS lines of %s (samples, executed, top stalls); total samples %d, instructions %d" % (pat, tot, sum(int(r[iexec] or 0) for r in data)))
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:25]:
            top = sorted([(h[i], int(r[i] or 0)) for i in sc], key=lambda kv: -kv[1])[:2]
            print("  %6s %10s  %-58s %s" % (r[isamp], r[iexec], r[isrc][:58], top))

main()
