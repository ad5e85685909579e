#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report.

    python tools/ncu_by_line.py <report.ncu-rep> <object.o> <kernel name, plain or mangled> [top N]

ncu's CSV source page is per SASS instruction; the line table comes from `nvdisasm -g` on the same object (built with
-lineinfo), matched by instruction offset.  Lines are reported as file:line with warp-level instructions executed, the share
of the kernel's total, and the stall samples attributed to them."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    table, cur, inside = {}, ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith("\t.section") or ln.startswith("//-----"):
            # a plain name (k_topk_fast) matches its Itanium-mangled section (.text._ZN3gdr11k_topk_fastE...): length-prefixed identifier
            key = kernel if kernel.startswith("_Z") else f"{len(kernel)}{kernel}E"
            inside = (key in ln) if ".text." in ln else False
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(obj, kernel)
    m = re.match(r"^_ZN3gdr(\d+)", kernel)
    short = kernel[m.end():m.end() + int(m.group(1))] if m else kernel          # the identifier inside a mangled name, or the plain name
    txt = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + os.environ.get("NCU_KERNEL", short), "--page", "source", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ia, ii, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    base = int(rows[hdr + 1][ia], 16)
    by_line, total, tot_samp = collections.defaultdict(lambda: [0, 0]), 0, 0
    for r in rows[hdr + 1:]:
        if len(r) <= ii or not r[ia].startswith("0x"):
            continue
        off = int(r[ia], 16) - base
        key = table.get(off, (("?", 0), ""))[0]
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        by_line[key][0] += n
        by_line[key][1] += s
        total += n
        tot_samp += s
    print(f"kernel {kernel}: {total} warp instructions, {tot_samp} stall samples")
    by_file = collections.defaultdict(lambda: [0, 0])
    for (f, l), (n, s) in by_line.items():
        by_file[f][0] += n
        by_file[f][1] += s
    for f, (n, s) in sorted(by_file.items(), key=lambda x: -x[1][0]):
        print(f"  {f:28s} {n:10d} inst {100.0 * n / total:5.1f}%   {s:7d} samples {100.0 * s / max(1, tot_samp):5.1f}%")
    print("top lines by instructions:")
    for (f, l), (n, s) in sorted(by_line.items(), key=lambda x: -x[1][0])[:top]:
        print(f"  {f}:{l:<5d} {n:10d} inst {100.0 * n / total:5.1f}%   {s:7d} samples {100.0 * s / max(1, tot_samp):5.1f}%")
    print("top lines by stall samples:")
    for (f, l), (n, s) in sorted(by_line.items(), key=lambda x: -x[1][1])[:top // 2]:
        print(f"  {f}:{l:<5d} {n:10d} inst {100.0 * n / total:5.1f}%   {s:7d} samples {100.0 * s / max(1, tot_samp):5.1f}%")


if __name__ == "__main__":
    main()
