"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.

    python tools/ncu_launch_shares.py gpurun_out/r1_launches_C2.csv profiles/r1_C2_launch_shares.csv [skip_launches]

Times under ncu are cold-cache and serialised; only the SHARE of each kernel is meaningful (DESIGN.md "Measurement").
`skip_launches` drops the first n launches (scene upload / warm-up noise) before aggregating.
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.split("(")[0]
    m = re.match(r"(gsb::\w+(<[^>]*>)?)", name)
    if m:
        return m.group(1)
    m = re.search(r"(\w+)(<|$)", name.replace("at::native::", "").replace("(anonymous namespace)::", ""))
    return "torch/cub: " + (m.group(1) if m else name)[:48]


def main(src, out, skip=0):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1 + skip:]:
        k = short(r[ki])
        t, n = tot.get(k, (0.0, 0))
        tot[k] = (t + float(r[vi].replace(",", "")), n + 1)
    allns = sum(t for t, _ in tot.values())
    ours = sum(t for k, (t, _) in tot.items() if k.startswith("gsb::"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_all_pct", "share_of_gsb_pct"])
        for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            w.writerow([k, n, f"{t / 1e3:.1f}", f"{t / 1e3 / n:.2f}", f"{100 * t / allns:.2f}",
                        f"{100 * t / ours:.2f}" if k.startswith("gsb::") and ours else ""])
    print("wrote", out, len(rows) - 1 - skip, "launches;", f"gsb kernels = {100 * ours / allns:.1f}% of captured GPU time")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
