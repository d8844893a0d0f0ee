"""Where does host time go in one forward+backward call?  (cProfile + wall clock; run on the GPU box.)"""
import cProfile
import math
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
from gs_b200 import synth
from diff_gaussian_rasterization import _C

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scene = synth.config_scene(cfg).to("cuda")
W, H = synth.config_image(cfg)
cam = synth.make_camera(W, H).to("cuda")
bg = torch.zeros(3, device="cuda")
dL = synth.grad_image(W, H, 5).cuda()
E = torch.Tensor([])
tx, ty = math.tan(cam.FoVx * .5), math.tan(cam.FoVy * .5)


def step():
    a = (bg, scene.means3D, E, scene.opacity, scene.scales, scene.rotations, 1.0, E, cam.world_view_transform,
         cam.full_proj_transform, tx, ty, H, W, scene.sh, scene.degrees, cam.camera_center, False, False)
    R, color, radii, gb, bb, ib = _C.rasterize_gaussians(*a)
    return _C.rasterize_gaussians_backward(bg, scene.means3D, radii, E, scene.scales, scene.rotations, 1.0, E, a[8], a[9], tx, ty, dL,
                                           scene.sh, scene.degrees, cam.camera_center, gb, R, bb, ib, 0.0, False)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{cfg}: host enqueue {1e3 * (t1 - t0) / 50:.3f} ms/step, incl. final drain {1e3 * (t2 - t0) / 50:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
