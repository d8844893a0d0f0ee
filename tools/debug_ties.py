import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "reduced-3dgs_b200"): sys.path.insert(0, os.path.join(ROOT, p))
import ours, gs_oracle
from gs_b200 import synth
g = torch.Generator().manual_seed(7)
P, W, H = 30_000, 128, 128
xyz = torch.rand(P, 3, generator=g) * 2 - 1
xyz[:, 2] = 0.0; xyz[::3, 2] = 0.25
scales = torch.full((P, 3), 0.02) * (0.5 + torch.rand(P, 3, generator=g))
q = torch.nn.functional.normalize(torch.randn(P, 4, generator=g))
op = torch.randn(P, 1, generator=g) - 2.0
sh = torch.randn(P, 1, 3, generator=g); deg = torch.zeros(P, 1, dtype=torch.int32)
scene = synth.Scene(xyz.contiguous(), op, scales.contiguous(), q.contiguous(), sh, deg)
cam = synth.make_camera(W, H); bg = torch.zeros(3)
args, out, fwd = ours.run_forward(scene, cam, bg)
kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center, W=W, H=H, tan_fovx=math.tan(cam.FoVx*.5), tan_fovy=math.tan(cam.FoVy*.5))
o = gs_oracle.forward(xyz, op, scales, q, sh, deg, bg=bg, **kw)
print("R", fwd["num_rendered"], o["num_rendered"])
for k in ("radii", "keys", "point_list", "ranges"):
    a, b = np.asarray(o[k]).reshape(-1), fwd[k].reshape(-1)
    bad = np.nonzero(a != b)[0]
    print(k, "mismatch", bad.size, "of", a.size, "first", bad[:5], a[bad[:5]], b[bad[:5]])
cnt = o["ranges"][:,1]-o["ranges"][:,0]
print("tile counts max", cnt.max(), "tiles>2048", (cnt>2048).sum(), ">8192", (cnt>8192).sum())
pl_o, pl_f = o["point_list"], fwd["point_list"]
bad = np.nonzero(pl_o != pl_f)[0]
if bad.size:
    tiles = np.searchsorted(o["ranges"][:,1], bad, side="right")
    print("bad tiles", np.unique(tiles)[:20], "counts", cnt[np.unique(tiles)[:20]])
    t = np.unique(tiles)[0]; r0, r1 = o["ranges"][t]
    print("segment sorted-set equal:", np.array_equal(np.sort(pl_o[r0:r1]), np.sort(pl_f[r0:r1])))
    print("oracle keys low32 in seg first 8", (o["keys"][r0:r0+8] & 0xffffffff), "ids o", pl_o[r0:r0+8], "ids f", pl_f[r0:r0+8])
