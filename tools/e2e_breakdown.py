"""Where does the end-to-end step go?  Variants of bench.py's e2e loop (config = argv[1], default C2), each timed with CUDA events over 20 steps."""
import math, os, sys, time
from types import SimpleNamespace
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
import bench
from gs_b200 import synth
from gaussian_renderer import render

dev = torch.device("cuda", 0)
CFG = sys.argv[1] if len(sys.argv) > 1 else "C2"
_name, W, H, scene, quant, _prune = bench.build_workload(SimpleNamespace(config=CFG, points=0), dev, 0, 1)
cams = [c.to(dev) for c in bench.bench_cameras(W, H, 4)]
G_host = synth.grad_image(W, H, 1000).pin_memory()
G_res = G_host.to(dev)
bg = torch.zeros(3, device=dev)
pipe = SimpleNamespace(debug=False, convert_SHs_python=False, compute_cov3D_python=False)
pc = bench.ModelView(scene, dev, None if quant is None else quant.to(dev), None)
cam_host = [torch.cat([c.world_view_transform.flatten(), c.full_proj_transform.flatten(), c.camera_center.flatten()]).cpu().pin_memory() for c in cams]
side = torch.cuda.Stream(device=dev); copy_done = torch.cuda.Event(); G_dev = torch.empty_like(G_host, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step(i, copy=True, item=True, set_none=True, host_t=None):
    v = i % 4
    t0 = time.perf_counter()
    cm = cam_host[v].to(dev, non_blocking=True)
    if copy:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            G_dev.copy_(G_host, non_blocking=True); copy_done.record(side)
        Gd = G_dev
    else:
        Gd = G_res
    cam = SimpleNamespace(FoVx=cams[v].FoVx, FoVy=cams[v].FoVy, image_height=H, image_width=W, world_view_transform=cm[:16].view(4, 4),
                          full_proj_transform=cm[16:32].view(4, 4), camera_center=cm[32:35])
    if set_none:
        for p in pc.params(): p.grad = None
        if pc.quant is not None: pc.quant.grads = None
    t1 = time.perf_counter()
    pkg = render(cam, pc, pipe, bg)
    t2 = time.perf_counter()
    if copy: torch.cuda.current_stream().wait_event(copy_done)
    loss = (pkg["render"] * Gd).sum()
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    r = float(loss.item()) if item else 0.0
    t5 = time.perf_counter()
    if host_t is not None: host_t.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4])
    return r

def run(name, **kw):
    for i in range(8): step(i, **kw)
    torch.cuda.synchronize()
    import gc; gc.collect(); gc.disable()
    ms = 0.0; ht = []
    for i in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i, host_t=ht, **kw); e1.record(); torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
    gc.enable()
    import numpy as np
    h = np.median(np.array(ht), axis=0) * 1e3
    print(f"{name:42s} {ms/20:.3f} ms/step   host ms [prep {h[0]:.3f} | render {h[1]:.3f} | loss {h[2]:.3f} | backward {h[3]:.3f} | item {h[4]:.3f}]", flush=True)

run("full e2e")
run("no dL copy", copy=False)
run("no dL copy, no item()", copy=False, item=False)
run("no dL copy, no item(), grads kept", copy=False, item=False, set_none=False)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
