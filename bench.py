#!/usr/bin/env python
"""bench.py — rendered Mpixels/s, forward+backward, of the splat-rasterizer hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # our CUDA path (default workload C3, the north-star target config)
    python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 # the unmodified reference rasterizer (oracle/_ref)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one view: forward (preprocess -> binning -> sort -> composite) + backward, on synthetic data of the
BASELINE.json configuration (default C3 = configs[2], the configuration the north-star target is quoted on: 3 M Gaussians,
mixed SH degrees, codebook-quantised attributes, 1920x1080; `--config C2` = configs[1], 500 k fp32 Gaussians).  Prints ONE JSON line.
With N > 1 ranks the views are sharded, and the timed region closes >= 3 view batches with the gradient all-reduce (its time is
part of `value` and reported per batch as `coll_ms`).

  value   whole-job Mpix/s with inputs resident in HBM; every step is timed with its own CUDA-event pair on the
          stream the kernels run on, L2 is flushed (256 MB write) between steps outside the event pairs
  e2e     the same metric through the reference-facing public API (gaussian_renderer.render + autograd) with the
          step's inputs (camera matrices, dL/dimage) copied from pinned host memory and the loss read back
  roofline  dominant kernel: SURVEY.md §8(d) algorithmic bytes / its CUDA-event time (measured live, in the timed
          region, by the library's per-kernel event pairs) against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the reference arithmetic, all host cores) on one view of the same workload
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
from gs_b200 import synth  # noqa: E402

EMPTY = torch.Tensor([])


# ------------------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--points", type=int, default=0, help="override the number of Gaussians (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (pynvml), from the main thread in the gap
    between two steps (after the L2 flush is launched, before the next step's start event is recorded).
    Two earlier designs perturbed the measurement and were dropped: polling with the nvidia-smi binary stalled kernel launches
    for milliseconds per query, and an NVML polling thread occasionally held a driver lock across a step's launches
    (one step in a few hundred took 30-100 ms)."""

    def __init__(self, index, enabled=True):
        self.index, self.sm, self.reasons, self.max_sm, self.nv = index, [], set(), None, None
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.names = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                          "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                          "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            self.sample(keep=False)                  # the first call of each query initialises driver state (20-200 ms)
        except Exception:
            self.nv = None

    def sample(self, keep=True):
        if self.nv is None:
            return
        try:
            c = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            return
        if keep:
            self.sm.append(c)
            for n, bit in self.names.items():
                if r & bit:
                    self.reasons.add(n)

    def result(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable or sampling disabled"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "where": "main thread, between steps of the timed region (GPU busy with the L2 flush)"}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    parts = [p for p in vis.split(",") if p.strip()]
    if local < len(parts) and parts[local].strip().isdigit():
        return int(parts[local])
    return local


def bind_to_gpu_numa_node(gpu_index: int):
    """N > 1: keep this rank's host threads and the pinned staging buffers it allocates next on the CPUs NVML reports as local to
    its GPU (what `numactl --cpunodebind` does for a torchrun job): the per-step host->device copy of the end-to-end arm then does
    not cross the socket interconnect.  Best effort: any failure leaves the affinity as it was.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus = local & os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def bench_cameras(W, H, n=16):
    """Views around the canonical camera: small yaw steps so every view sees the whole cloud."""
    cams = []
    for i in range(n):
        th = math.radians((i - n / 2) * 1.5)
        Rc2w = np.array([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
        C = Rc2w @ np.array([0.0, 0.0, -4.0])
        cams.append(synth.make_camera(W, H, Rc2w, -Rc2w.T @ C))
    return cams


from gs_b200.model import GaussianModelView as ModelView  # noqa: E402  (the attributes gaussian_renderer.render reads)


def algorithmic_bytes(P, V, R, sumK, Npx, Nt, rho_f, rho_b, quant, mask):
    """SURVEY.md §8(d) formulas (bytes per frame)."""
    if quant:
        pre_in = P * (12 + 11 + 4) + 3 * max(sumK - V, 0) + 20 * 1024
    else:
        pre_in = P * 48 + 12 * sumK
    B_pre = pre_in + (P if mask else 0) + V * 40 + P * 8
    B_bin = P * 8 + P * 4 + V * 16 + R * 12 + R * 24 + R * 4 + Nt * 8
    B_rend = R * 40 * rho_f + Npx * 20
    B_bwdR = R * 40 * rho_b + Npx * 20 + V * 36 * 2
    B_bwdP = V * (12 + 24 + 12 + 16 + 16 + 3) + 12 * sumK * 2 + V * (12 + 12 + 16 + 4) + P * 12
    return dict(preprocess=B_pre, binning=B_bin, render_forward=B_rend, render_backward=B_bwdR, preprocess_backward=B_bwdP)


# ------------------------------------------------------------------------------------------------------
def build_workload(args, dev, rank, world):
    name = args.config
    W, H = synth.config_image(name)
    scene = synth.config_scene(name, args.points or None)
    quant = prune = None
    if name in ("C3", "C4", "C5"):
        quant = synth.quantise_scene(scene, seed=0)
        # the fp32 tensors the reference sees: its load_ply de-quantises ON THE GPU (gaussian_model.py:371-387 + get_scaling /
        # get_rotation), and CUDA exp / normalize differ from the CPU ones in the last ulp (a handful of radii out of 6 M)
        dq = quant.to(dev).dequantise()
        scene = synth.Scene(*[getattr(dq, f).cpu() for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees")])
    if name == "C4":
        prune = synth.prune_mask(scene.P, 4)
    return name, W, H, scene, quant, prune


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.impl == "reference" and rank != 0:
        return 0                                    # the reference is single-GPU: rank 0 alone runs it
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(physical_gpu_index(local)) if (world > 1 and args.impl == "ours") else None
    name, W, H, scene, quant, prune = build_workload(args, dev, rank, world)
    Npx = W * H
    if name == "C4":        # BASELINE.json configs[3]: the 64-view orbit batch, sharded over the ranks
        cams = [c.to(dev) for c in synth.orbit_cameras(64, W, H)]
    else:
        cams = [c.to(dev) for c in bench_cameras(W, H, 4 * world)]
    my_views = list(range(rank, len(cams), world)) if args.impl == "ours" else list(range(len(cams)))
    bg = torch.zeros(3, device=dev)
    G_host = synth.grad_image(W, H, 1000 + rank).pin_memory()
    G = G_host.to(dev)
    tanx = [math.tan(c.FoVx * 0.5) for c in cams]
    tany = [math.tan(c.FoVy * 0.5) for c in cams]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    K, Wm = args.steps, args.warmup

    sd = scene.to(dev)
    prune_d = None if prune is None else prune.to(dev)
    refC = None
    if args.impl == "ours":
        from diff_gaussian_rasterization import _C
        from gs_b200 import lib as gsl
        from gs_b200 import multi
        qd = None if quant is None else quant.to(dev)
        if world > 1:      # replicate the scene from rank 0 (the only model-state transfer of the sharded path)
            multi.broadcast_scene([sd.means3D, sd.opacity, sd.scales, sd.rotations, sd.sh])
        bands = None
        if name in ("C3", "C4", "C5"):                 # degree-banded models (Gaussians ordered by degree): only active SH coefficients travel
            bands = [int((sd.degrees.view(-1) == d).sum()) for d in range(4)]
        acc = multi.GradAccumulator(sd.P, 16 if qd is not None else sd.sh.shape[1], dev, band_counts=bands) if world > 1 else None

        def step(i, dL):
            v = my_views[i % len(my_views)]
            c = cams[v]
            if qd is not None:
                a = (bg, sd.means3D, EMPTY, EMPTY, EMPTY, EMPTY, 1.0, EMPTY, c.world_view_transform, c.full_proj_transform,
                     tanx[v], tany[v], H, W, EMPTY, sd.degrees, c.camera_center, False, False)
            else:
                a = (bg, sd.means3D, EMPTY, sd.opacity, sd.scales, sd.rotations, 1.0, EMPTY, c.world_view_transform,
                     c.full_proj_transform, tanx[v], tany[v], H, W, sd.sh, sd.degrees, c.camera_center, False, False)
            R, color, radii, gb, bb, ib = _C.rasterize_gaussians(*a, prune_mask=prune_d, quant=qd)
            grads = _C.rasterize_gaussians_backward(bg, sd.means3D, radii, EMPTY, a[4], a[5], 1.0, EMPTY, a[8], a[9], a[10], a[11],
                                                    dL, a[14], sd.degrees, a[16], gb, R, bb, ib, 0.0, False, prune_mask=prune_d,
                                                    quant=qd, accumulate_into=None if acc is None else acc.buffers(),
                                                    view_means2D=None if acc is None else acc.view_means2D)
            return R, color, radii, ib, grads
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        refC = build_ref.load()
        if refC is None:
            return reference_cpu_port(args, name, W, H, scene, cams, tanx, tany)
        ref_scene = sd.compact(~prune_d.bool()) if prune_d is not None else sd   # reference semantics of pruning: delete rows

        def step(i, dL):
            v = my_views[i % len(my_views)]
            c = cams[v]
            s = ref_scene
            a = (bg, s.means3D, EMPTY, s.opacity, s.scales, s.rotations, 1.0, EMPTY, c.world_view_transform, c.full_proj_transform,
                 tanx[v], tany[v], H, W, s.sh, s.degrees, c.camera_center, False, False)
            R, color, radii, gb, bb, ib = refC.rasterize_gaussians(*a)
            grads = refC.rasterize_gaussians_backward(bg, s.means3D, radii, EMPTY, s.scales, s.rotations, 1.0, EMPTY, a[8], a[9],
                                                      a[10], a[11], dL, s.sh, s.degrees, a[16], gb, R, bb, ib, 0.0, False)
            return R, color, radii, ib, grads

    def barrier():
        torch.cuda.synchronize()
        if world > 1 and args.impl == "ours":
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident measurement (`value`) ----------------
    # warm-up runs the SAME loop body as the timed region (L2 flush, event pair, step, batch-closing all-reduce) so that every lazily
    # initialised piece (kernel modules, caching-allocator blocks, event pools, NCCL channels) exists before timing starts.
    # --warmup is honoured as given (the contract's minimum of 3 applies).
    n_warm = max(Wm, 3)
    if name == "C4":
        # the 64-camera orbit's instance count varies 3x between front and side views: the steady state of a loop over such a view set
        # is reached after one pass (allocator high-water marks, binning capacity), so the warm-up covers this rank's views once
        n_warm = max(n_warm, len(my_views))
    multi_gpu = world > 1 and args.impl == "ours"
    n_batches = min(3, K) if multi_gpu else 1          # N > 1: the timed region closes >= 3 view batches with the gradient all-reduce

    def closes_batch(i, n, nb):
        return (i + 1) * nb // n != i * nb // n        # step i is the last of its batch (n steps split into nb batches)

    checked = []

    def close_batch():
        if not checked:                                # once, in the warm-up: the banded payload's premise holds for this scene
            assert acc.inactive_sh_is_zero(), "gradient outside the active SH bands"
            checked.append(True)
        acc.all_reduce()                               # SUM over 62 floats/Gaussian + 2 statistics, MAX over the radii (gs_b200/multi.py)
        acc.zero_()                                    # the next batch accumulates from zero (part of the batch's cost)
    sampler = ClockSampler(physical_gpu_index(local), enabled=(rank == 0 and not os.environ.get("GS_BENCH_NO_CLOCKS")))
    if args.impl == "ours":
        gsl.profile_enable(True)                 # per-kernel event pairs are part of the measured configuration: warm them up too
    for i in range(n_warm):
        flush.zero_()
        sampler.sample(keep=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step(i, G)
        e1.record()
        if multi_gpu and closes_batch(i, n_warm, 2):
            close_batch()
        if os.environ.get("GS_BENCH_DEBUG"):
            st_ = torch.cuda.memory_stats(dev)
            sys.stderr.write(f"  warm {i}: device allocs {st_['num_device_alloc']}, reserved {st_['reserved_bytes.all.current'] / 1e6:.0f} MB, active {st_['active_bytes.all.current'] / 1e6:.0f} MB, R={out[0]}\n")
    torch.cuda.synchronize()
    if args.impl == "ours":
        gsl.profile_read()                       # discard warm-up samples (recycles the events)
        launches0 = gsl.launch_count()
    import gc
    gc.collect()
    gc.disable()                                  # a cyclic-GC pause between two launches would show up as GPU idle time
    ms0 = torch.cuda.memory_stats(dev)
    barrier()
    host_t = []
    evs = []
    coll_evs = []
    for i in range(K):
        flush.zero_()
        sampler.sample()                         # in the gap: the step's launches below never overlap an NVML call
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_h = time.perf_counter()
        a_h = torch.cuda.memory_stats(dev)["num_device_alloc"] if os.environ.get("GS_BENCH_DEBUG") else 0
        out = step(i, G)
        if os.environ.get("GS_BENCH_DEBUG"):
            st_ = torch.cuda.memory_stats(dev)
            sys.stderr.write(f"  step {i}: new device allocs {st_['num_device_alloc'] - a_h}, reserved {st_['reserved_bytes.all.current'] / 1e6:.0f} MB, active {st_['active_bytes.all.current'] / 1e6:.0f} MB, R={out[0]}\n")
        host_t.append(time.perf_counter() - t_h)
        e1.record()
        evs.append((e0, e1))
        if multi_gpu and closes_batch(i, K, n_batches):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            close_batch()                        # one gradient all-reduce closes the view batch
            c1.record()
            coll_evs.append((c0, c1))
    barrier()
    gc.enable()
    coll_list = [a.elapsed_time(b) for a, b in coll_evs]
    coll_ms = sum(coll_list)
    R0, color0, radii0, ib0, _ = out             # (taken after the loop: holding a warm-up generation would grow the live set mid-run)
    clocks = sampler.result() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]
    if os.environ.get("GS_BENCH_DEBUG"):
        ms1 = torch.cuda.memory_stats(dev)
        sys.stderr.write("step_ms " + " ".join(f"{x:.2f}" for x in step_ms) + "\nhost_ms " + " ".join(f"{1e3 * x:.2f}" for x in host_t) +
                         f"\nalloc_retries {ms1['num_alloc_retries'] - ms0['num_alloc_retries']} device_allocs {ms1['num_device_alloc'] - ms0['num_device_alloc']}"
                         f" device_frees {ms1['num_device_free'] - ms0['num_device_free']}\n")
    total_ms = sum(step_ms) + coll_ms
    if multi_gpu:
        t = torch.tensor([total_ms, coll_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, coll_max = float(t[0].item()), float(t[1].item())
    n_gpus = world if args.impl == "ours" else 1
    value = n_gpus * K * Npx / (total_ms * 1e-3) / 1e6
    prof, launches = {}, None
    if args.impl == "ours":
        prof = gsl.profile_read()
        gsl.profile_enable(False)
        launches = gsl.launch_count() - launches0

    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, dev, scene, quant, prune, cams, tanx, tany, my_views, G_host, H, W, K, flush, refC, world, n_gpus, rank,
                      n_batches, closes_batch)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = peaks()
    radii_t = radii0
    V = int((radii_t > 0).sum())
    deg = sd.degrees.view(-1)[: radii_t.shape[0]] if args.impl == "ours" else None
    roof = None
    if args.impl == "ours":
        from diff_gaussian_rasterization import _C as _C2
        sumK = int((((deg.long() + 1) ** 2)[radii_t > 0]).sum())
        st = _C2.export_state(None, None, ib0, 0, W, H)
        ncon = st["n_contrib"].float()[None, None]
        tile_max = torch.nn.functional.max_pool2d(ncon, 16, ceil_mode=True)
        rho = float(tile_max.sum().item()) / max(R0, 1)
        Nt = ((W + 15) // 16) * ((H + 15) // 16)
        B = algorithmic_bytes(sd.P, V, R0, sumK, Npx, Nt, rho, rho, quant is not None, prune is not None)
        groups = {"preprocess": ["preprocess"], "binning": ["tile_scan", "scatter", "tile_sort", "tile_sort_large"],
                  "render_forward": ["render_forward"], "render_backward": ["render_backward"], "preprocess_backward": ["preprocess_backward"]}
        per_kernel = {k: {"ms_per_step": v[0] / K, "launches_per_step": v[1] / K} for k, v in prof.items()}
        stage_ms = {g: sum(prof.get(k, (0, 0))[0] for k in ks) / K for g, ks in groups.items()}
        dom = max(["render_forward", "render_backward", "preprocess", "preprocess_backward"], key=lambda g: stage_ms[g])
        ach = B[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        frame_bytes = sum(B.values())
        kern_ms = sum(stage_ms.values())
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this config (profiles/traffic.json,
        # written by tools/ncu_summary.py); null when this config / kernel has no capture
        traffic = None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
        if os.path.isfile(tpath) and not args.points:
            traffic = json.load(open(tpath)).get(name, {}).get(dom)
        fwd_ms = stage_ms["preprocess"] + stage_ms["binning"] + stage_ms["render_forward"]
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic, "algorithmic_bytes": int(B[dom]), "kernel_ms": round(stage_ms[dom], 4),
                "stages": {g: {"ms": round(stage_ms[g], 4), "alg_MB": round(B[g] / 1e6, 1),
                               "GBps": round(B[g] / max(stage_ms[g], 1e-9) / 1e6, 1),
                               "frac": round(B[g] / max(stage_ms[g], 1e-9) / 1e6 / peak, 4)} for g in groups},
                "frame": {"alg_MB": round(frame_bytes / 1e6, 1), "kernel_ms": round(kern_ms, 4),
                          "frac": round(frame_bytes / max(kern_ms, 1e-9) / 1e6 / peak, 4)},
                "fwd_only": {"kernel_ms": round(fwd_ms, 4), "Mpix_s": round(Npx / max(fwd_ms, 1e-9) / 1e3, 1),
                             "note": "sum of the forward kernels' device time (preprocess + binning + render_forward)"},
                "rho_list_consumed": round(rho, 4), "kernels": per_kernel,
                "timing": "per-kernel CUDA-event pairs recorded by the library on the launch stream INSIDE the timed region (their cost is part of `value`); "
                          "render_backward's pair includes the memset of its accumulator, scatter's the pre-write of the bucket array"}

    # ---------------- CPU baseline: the oracle on the host cores, one view of the same workload ----------------
    cpu = None
    if args.impl == "ours" and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(scene, prune, [c.to("cpu") for c in cams], W, H, n_views={"C1": 8, "C2": 6}.get(name, 2))

    line = {"metric": "rendered Mpixels/s fwd+bwd", "value": round(value, 2), "unit": "Mpix/s", "n_gpus": n_gpus, "steps": K,
            "warmup": n_warm, "ms_per_step": round(total_ms / K, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {scene.P} Gaussians, {W}x{H}, fwd+bwd, one view per step"
                                   + (", codebook-quantised (fused dequant)" if (quant is not None and args.impl == 'ours') else "")
                                   + (", prune mask" if prune is not None else ""),
                       "points": scene.P, "image": [W, H], "visible": V, "instances_R": int(R0),
                       "l2": "256 MB flush write between steps (outside the per-step event pairs)",
                       "views": ("64-camera orbit (SURVEY §8(d) C4)" if name == "C4" else f"{len(cams)} cameras, +-1.5 deg yaw steps around the canonical one"),
                       "parallelism": f"views sharded over {n_gpus} GPU(s), scene replicated"
                                      + (f", {n_batches} view batches in the timed region, each closed by one gradient all-reduce" if n_gpus > 1 else "")
                                      + (f", rank 0 bound to its GPU's {numa_cpus} local CPUs" if numa_cpus else "")},
            "impl": args.impl, "clocks": clocks,
            "step_ms": {"min": round(min(step_ms), 4), "median": round(float(np.median(step_ms)), 4), "max": round(max(step_ms), 4)}}
    if multi_gpu:
        line["collective"] = {"batches": n_batches, "coll_ms_per_batch": [round(x, 4) for x in coll_list], "coll_ms_max_over_ranks_total": round(coll_max, 4),
                              "payload_MB": round(acc.payload_floats * 4 / 1e6, 1), "floats_per_gaussian": acc.floats_per_gaussian,
                              "dense_payload_MB": round(sd.P * (acc.floats_per_gaussian + 2) * 4 / 1e6, 1),
                              "busbw_GBps": round(2 * (world - 1) / world * acc.payload_floats * 4 / (min(coll_list) * 1e-3) / 1e9, 1) if coll_list else None,
                              "what": "per batch: all_reduce(SUM) of the 62 floats/Gaussian the optimiser consumes + 2 statistics (degree-banded models: only "
                                      "the ACTIVE SH coefficients of each degree group travel), all_reduce(MAX) of the radii, pack / unpack and "
                                      "re-zeroing of the accumulators; warmed twice in the warm-up loop; included in `value`"}
    if e2e is not None:
        line["e2e"] = e2e
    if launches is not None:
        line["gpu_launches"] = int(launches)
    if roof is not None:
        line["roofline"] = roof
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if args.impl == "reference":
        line["cpu_baseline"] = {"value": line["value"], "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                "sample": "the reference's own implementation of this path is CUDA (no CPU rasterizer exists): "
                                          "oracle/_ref/_refC.so = unmodified reference sources + GLM stand-in, timed on the same B200, "
                                          "debug syncs off (the faster of the two ways the reference can run)"}
    print(json.dumps(line), flush=True)
    if world > 1 and args.impl == "ours":
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, dev, scene, quant, prune, cams, tanx, tany, my_views, G_host, H, W, K, flush, refC, world, n_gpus, rank,
            n_batches=1, closes_batch=None):
    """The call a user makes: render(camera, model, pipe, bg) + loss.backward(), inputs from pinned host memory.
    N > 1: gradients accumulate in the parameters' .grad over a view batch and every batch ends with the all-reduce of those
    gradients (what a data-parallel training loop on this API does), inside the timed region."""
    import torch.distributed as dist
    Npx = W * H
    cam_host = [torch.cat([c.world_view_transform.flatten(), c.full_proj_transform.flatten(), c.camera_center.flatten()]).cpu().pin_memory()
                for c in cams]
    bg = torch.zeros(3, device=dev)
    h2d = G_host.numel() * 4 + cam_host[0].numel() * 4
    # both arms use the same user-level pipelining: the 24.9 MB dL/dimage copy runs on a side stream while the forward renders
    side = torch.cuda.Stream(device=dev)
    copy_done = torch.cuda.Event()
    G_dev = torch.empty_like(G_host, device=dev)

    step_start = torch.cuda.Event()

    def h2d_side(src, after_current=True):
        # the copy may not start before the previous step has consumed G_dev and the timed region has begun (`step_start`, recorded
        # on the main stream at the top of every step); it does not wait for the forward that was just enqueued
        if after_current:
            side.wait_stream(torch.cuda.current_stream())
        else:
            side.wait_event(step_start)
        with torch.cuda.stream(side):
            G_dev.copy_(src, non_blocking=True)
            copy_done.record(side)
        return G_dev
    if args.impl == "ours":
        from gaussian_renderer import render
        pipe = SimpleNamespace(debug=False, convert_SHs_python=False, compute_cov3D_python=False)
        pc = ModelView(scene, dev, None if quant is None else quant.to(dev), None if prune is None else prune.to(dev))

        def one(i):
            v = my_views[i % len(my_views)]
            step_start.record()
            cm = cam_host[v].to(dev, non_blocking=True)
            cam = SimpleNamespace(FoVx=cams[v].FoVx, FoVy=cams[v].FoVy, image_height=H, image_width=W,
                                  world_view_transform=cm[:16].view(4, 4), full_proj_transform=cm[16:32].view(4, 4),
                                  camera_center=cm[32:35])
            if world == 1:
                for p in pc.params():
                    p.grad = None
                if pc.quant is not None:
                    pc.quant.grads = None
            pkg = render(cam, pc, pipe, bg)
            # dL/dimage is only needed by the backward: its 24.9 MB copy is enqueued on a side stream once the forward is launched and
            # runs under it (both arms; enqueueing it first delays the forward's first kernel by the copy call's host time)
            Gd = h2d_side(G_host, after_current=False)
            torch.cuda.current_stream().wait_event(copy_done)
            loss = (pkg["render"] * Gd).sum()
            loss.backward()
            return float(loss.item())

        def close_batch_e2e():
            gl = [p.grad for p in pc.params() if p.grad is not None]
            if pc.quant is not None and getattr(pc.quant, "grads", None):
                gl += [g for g in pc.quant.grads.values() if g is not None and g.numel()]
            for g in gl:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
            for p in pc.params():
                p.grad = None
            if pc.quant is not None:
                pc.quant.grads = None
    else:
        sd = scene.to(dev)
        if prune is not None:
            sd = sd.compact(~prune.to(dev).bool())

        def one(i):
            v = my_views[i % len(my_views)]
            step_start.record()
            cm = cam_host[v].to(dev, non_blocking=True)
            a = (bg, sd.means3D, EMPTY, sd.opacity, sd.scales, sd.rotations, 1.0, EMPTY, cm[:16].view(4, 4).contiguous(),
                 cm[16:32].view(4, 4).contiguous(), tanx[v], tany[v], H, W, sd.sh, sd.degrees, cm[32:35].contiguous(), False, False)
            R, color, radii, gb, bb, ib = refC.rasterize_gaussians(*a)
            Gd = h2d_side(G_host, after_current=False)
            torch.cuda.current_stream().wait_event(copy_done)
            loss = (color * Gd).sum()
            refC.rasterize_gaussians_backward(bg, sd.means3D, radii, EMPTY, sd.scales, sd.rotations, 1.0, EMPTY, a[8], a[9], a[10],
                                              a[11], Gd, sd.sh, sd.degrees, a[16], gb, R, bb, ib, 0.0, False)
            return float(loss.item())
    multi_gpu = world > 1 and args.impl == "ours"
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        one(i)
        if multi_gpu and closes_batch(i, n_warm, 2):
            close_batch_e2e()
    torch.cuda.synchronize()
    if multi_gpu:
        dist.barrier()
    ms = 0.0
    close_ms = []
    import gc
    gc.collect()
    gc.disable()
    for i in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one(i)
        closing = multi_gpu and closes_batch(i, K, n_batches)
        if closing:
            ec = torch.cuda.Event(enable_timing=True)
            ec.record()
            close_batch_e2e()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        if closing:
            close_ms.append(ec.elapsed_time(e1))
    gc.enable()
    if world > 1 and args.impl == "ours":
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": round(n_gpus * K * Npx / (ms * 1e-3) / 1e6, 2), "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 4, "ms_per_step": round(ms / K, 4),
            **({"close_ms_per_batch": [round(x, 3) for x in close_ms]} if close_ms else {}),
            "api": ("gaussian_renderer.render + loss.backward()" + (f"; {n_batches} view batches, each closed by all_reduce of the parameter gradients" if multi_gpu else ""))
                   if args.impl == "ours" else "_C.rasterize_gaussians + _C.rasterize_gaussians_backward"}


def cpu_baseline(scene, prune, cams, W, H, n_views=2, budget_s=60.0):
    """Oracle (CPU port of the reference arithmetic) forward+backward on a bounded sample of the same workload: a FIXED number
    of whole views per configuration (so the figure is comparable between boxes; ~10-30 s of CPU work on the box's cores,
    OpenMP), cut short only if it exceeds budget_s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gs_oracle
    bg = np.zeros(3, np.float32)
    dL = synth.grad_image(W, H, 1000).numpy()
    t_f = t_b = 0.0
    n = 0
    for cam in cams[:n_views]:
        kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center, W=W, H=H,
                  tan_fovx=math.tan(cam.FoVx * 0.5), tan_fovy=math.tan(cam.FoVy * 0.5))
        t0 = time.time()
        fwd = gs_oracle.forward(scene.means3D, scene.opacity, scene.scales, scene.rotations, scene.sh, scene.degrees, bg=bg,
                                prune_mask=None if prune is None else prune.numpy(), **kw)
        t1 = time.time()
        if prune is None:
            gs_oracle.backward(fwd, dL, scene.means3D, scene.scales, scene.rotations, scene.sh, scene.degrees, bg=bg, **kw)
        t2 = time.time()
        t_f, t_b, n = t_f + (t1 - t0), t_b + (t2 - t1), n + 1
        if t_f + t_b >= budget_s:
            break
    return {"value": round(n * W * H / (t_f + t_b) / 1e6, 4), "unit": "Mpix/s", "cores": gs_oracle.num_threads(), "kind": "port",
            "sample": f"{n} view(s) of the same workload, forward {t_f:.2f} s + backward {t_b:.2f} s in total (OpenMP, {os.cpu_count()} host CPUs)"}


def reference_cpu_port(args, name, W, H, scene, cams, tanx, tany):
    """--impl reference when oracle/_ref/_refC.so is not available: time the CPU oracle port instead."""
    cpu = cpu_baseline(scene, None, [c.to("cpu") for c in cams], W, H, n_views={"C1": 8, "C2": 6}.get(name, 2))
    line = {"metric": "rendered Mpixels/s fwd+bwd", "value": cpu["value"], "unit": "Mpix/s", "n_gpus": 1, "steps": 1, "warmup": 0,
            "ms_per_step": round(W * H / cpu["value"] / 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": f"{name}: {scene.P} Gaussians, {W}x{H}, fwd+bwd"},
            "impl": "reference", "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
