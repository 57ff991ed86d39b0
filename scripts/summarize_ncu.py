#!/usr/bin/env python
"""Turn gpurun_out/prof_full.ncu-rep (ncu --set full) + launches.csv into a markdown summary
under profiles/.  Usage: summarize_ncu.py <tag> [note]"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
]


def main():
    tag = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rep = os.path.join(OUT, "prof_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    lines = ["# ncu summary %s" % tag, "", note, "",
             "Source: `ncu --set full --clock-control none` on `python bench.py --steps 3 --warmup 3` (1 x B200), one "
             "capture per kernel of one forward+backward step at B=32 N=8000 V=64 K=21.  ncu serialises kernels and "
             "flushes caches between replays, so times are cold-cache upper bounds; compare SHARES with `stages_ms` "
             "of the bench line, not absolutes.", ""]
    for r in rows[2:]:
        name = r[ki]
        if "dpc_" not in name:
            continue
        lines.append("## `%s`" % name[:90])
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("| %s (`%s`) | %s %s |" % (label, k, r[i], units[i]))
        st = sorted(((float(r[hdr.index(h)]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for h in stalls), reverse=True)[:6]
        lines.append("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st))
        lines.append("")
    lp = os.path.join(OUT, "launches.csv")
    if os.path.isfile(lp):
        rws = [r for r in csv.reader(open(lp)) if len(r) > 5 and r[0].isdigit()]
        agg = collections.defaultdict(list)
        for r in rws:
            agg[r[4][:70]].append(float(r[-1]))
        lines += ["## launch list (gpu__time_duration, all launches of the bench command)", "", "| kernel | launches | avg us | min us |", "|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append("| `%s` | %d | %.2f | %.2f |" % (k, len(v), sum(v) / len(v) / 1000, min(v) / 1000))
        lines.append("")
    for f in ("bench.json", "bench_noflush.json"):
        fp = os.path.join(OUT, f)
        if os.path.isfile(fp):
            lines += ["## %s" % f, "", "```", open(fp).read().strip(), "```", ""]
    path = os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag)
    open(path, "w").write("\n".join(lines))
    print("wrote", path)


if __name__ == "__main__":
    main()
