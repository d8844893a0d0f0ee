#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the mnemonics that prove which hardware paths the shipped
libgs_b200.so uses (TMA bulk copies UBLKCP / tensor-map TMA UTMALDG, mbarrier SYNCS, tensor-core HMMA / UTC*MMA, packed fp32
FFMA2, vector reductions REDG, shared-memory atomics ATOMS, MUFU), plus registers / shared memory per kernel from the cubin.
Runs without a GPU:  python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "reduced-3dgs_b200", "gs_b200", "libgs_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "REDG", "RED", "ATOMS", "ATOMG",
         "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "VOTE", "SHFL", "MATCH"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True)
    usage = {}
    cur = None
    for ln in (res.stdout + res.stderr).split("\n"):
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            usage[cur] = " ".join(re.findall(r"(REG:\d+|SHARED:\d+|STACK:\d+)", ln))
            cur = None
    kernels = collections.OrderedDict()
    name = None
    for ln in sass.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and name:
            op = m.group(1)
            kernels[name]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[name][w] += 1
    dm = demangle(list(kernels))
    print("# SASS summary of", os.path.relpath(SO, ROOT), "(cuobjdump -sass; static instruction counts per kernel)")
    print("# columns: total instructions | resources | non-zero counts of the watched mnemonics")
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", dm.get(k, k)).replace("void ", "")
        watched = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short:60s} {c['total']:6d} | {usage.get(k, ''):28s} | {watched}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library totals:", " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))
    print("# absent (count 0 in every kernel):", " ".join(w for w in WATCH if not tot[w]))


if __name__ == "__main__":
    sys.exit(main())
