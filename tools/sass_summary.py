#!/usr/bin/env python
"""Mnemonic count per kernel of the built library (cuobjdump -sass, no GPU needed) -> profiles/rNN_sass_summary.txt, and the
tcgen05 kernel's SASS in full -> profiles/rNN_k_score_umma.sass.   usage: python tools/sass_summary.py [round-tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gdr_b200", "lib", "libgdr_b200.so")
WATCH = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "ACQBULK", "ELECT", "LDG", "STG", "LDS", "STS",
         "ATOMS", "ATOMG", "REDG", "FFMA", "SHFL", "MUFU"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per, name, umma = collections.OrderedDict(), None, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and name:
            per[name]["_n"] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    per[name][w] += 1
        if name and "k_score_umma" in name:
            umma.append(line)
    out = ["# SASS mnemonic summary of gdr_b200/lib/libgdr_b200.so (cuobjdump -sass, sm_100a), " + tag,
           "# UTCHMMA = tcgen05.mma, UTMALDG = TMA cp.async.bulk.tensor, UBLKCP = cp.async.bulk (tile records), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,",
           "# UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, LDGSTS = cp.async, ACQBULK = griddepcontrol.wait (programmatic dependent launch)"]
    for fn, c in per.items():
        out.append(fn)
        out.append("    instructions=%d %s" % (c["_n"], {k: v for k, v in c.items() if k != "_n"}))
    open(os.path.join(ROOT, "profiles", tag + "_sass_summary.txt"), "w").write("\n".join(out) + "\n")
    open(os.path.join(ROOT, "profiles", tag + "_k_score_umma.sass"), "w").write("\n".join(umma) + "\n")
    print("kernels:", len(per), "umma lines:", len(umma))


if __name__ == "__main__":
    main()
