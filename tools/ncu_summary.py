"""Condense an ncu report (.ncu-rep, read here without a GPU) into the per-kernel summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_C3_kernels.csv
"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
        ("smsp__inst_executed_op_global_red.sum", "global_red_inst"), ("smsp__inst_executed_op_shared_atom.sum", "shared_atom_inst")]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3 or "Kernel Name" not in rows[0]:
        sys.exit(f"{rep}: no kernels in the report (missing or empty capture) - {out} left untouched")
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([n + (f" [{units[i]}]" if units[i] else "") for i, n in idx])
        for r in rows[2:]:
            w.writerow([r[i][:70] if n == "kernel" else r[i] for i, n in idx])
    print("wrote", out, len(rows) - 2, "kernels")
    return rows


def traffic(rows):
    """{kernel id: dram bytes read+written per launch} — feeds bench.py's roofline.traffic (profiles/traffic.json)."""
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ids = [("render_forward", "render_forward"), ("render_backward", "render_backward"), ("preprocess_backward", "preprocess_backward"),
           ("preprocess_kernel", "preprocess"), ("scatter", "scatter"), ("tile_sort_dist", "tile_sort"), ("tile_prefix", "tile_prefix"),
           ("tile_scan", "tile_scan")]
    out = {}
    for r in rows[2:]:
        for pat, kid in ids:
            if pat in r[ki]:
                b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
                out[kid] = int(b)
                break
    return out


if __name__ == "__main__":
    rows = main(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 4:          # ... <traffic.json> <config key>: merge this capture's DRAM traffic per launch
        import json, os
        path, key = sys.argv[3], sys.argv[4]
        d = json.load(open(path)) if os.path.isfile(path) else {}
        d[key] = traffic(rows)
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
        print("updated", path, key, d[key])
