#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few metrics the roofline discussion uses.
usage: python tools/ncu_summary.py <report.ncu-rep> [more ...]  -> text on stdout (committed under profiles/)."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max",
]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"{rep}: empty")
            continue
        h, units = rows[0], rows[1]
        print(f"== {rep}")
        for r in rows[2:]:
            print(f"-- {r[h.index('Kernel Name')]}")
            for m in METRICS:
                if m in h:
                    i = h.index(m)
                    print(f"   {m:68s} {r[i]:>16s} {units[i]}")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        heads = [i for i, r in enumerate(srows) if "Source" in r and "# Samples" in r]
        for hi_, start in enumerate(heads):
            hh = srows[start]
            end = heads[hi_ + 1] - 1 if hi_ + 1 < len(heads) else len(srows)
            body = [r for r in srows[start + 1:end] if len(r) > hh.index("# Samples")]
            stall = {c: 0.0 for c in hh if c.startswith("stall_") and "Not" not in c}
            for r in body:
                for c in stall:
                    v = r[hh.index(c)]
                    if v not in ("", "0"):
                        stall[c] += float(v)
            tot = sum(stall.values()) or 1.0
            top = sorted(stall.items(), key=lambda x: -x[1])[:6]
            print(f"   warp-state samples, launch {hi_}: " + ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in top))


if __name__ == "__main__":
    main()
