"""GPU probe: which fp32 formulas reproduce torch's CUDA exp() and F.normalize() bit for bit?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
from gs_b200 import synth
from diff_gaussian_rasterization import _C

scene = synth.make_scene(200_000, 99, mixed_degrees=True)
q = synth.quantise_scene(scene).to("cuda")
d = q.dequantise()
s, r = _C.debug_dequant(q)
torch.cuda.synchronize()
print("scales: mismatches vs torch.exp   ", int((s != d.scales).sum()), "/", s.numel(), "max ulp", float(((s - d.scales).abs() / torch.abs(d.scales) / 1.19e-7).max()))
print("rot   : mismatches vs F.normalize ", int((r != d.rotations).sum()), "/", r.numel())
# candidates for the norm, evaluated exactly on the CPU
c = q.centers.cpu()
g = torch.cat([c[18][q.ids_rot[:, 0:1].long().cpu()], c[19][q.ids_rot[:, 1:].long().cpu()]], dim=1).numpy()
t = d.rotations.cpu().numpy()
f32, f64 = np.float32, np.float64
def fma(a, b, c): return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)
a, b, cc, dd = g[:, 0], g[:, 1], g[:, 2], g[:, 3]
cands = {
  "seq_fma (r*r, fma x, fma y, fma z)": fma(dd, dd, fma(cc, cc, fma(b, b, a * a))),
  "seq_nofma ((r2+x2)+y2)+z2": ((a * a + b * b) + cc * cc) + dd * dd,
  "pairwise (r2+x2)+(y2+z2)": (a * a + b * b) + (cc * cc + dd * dd),
  "pairwise_fma fma(x,x,r2)+fma(z,z,y2)": fma(b, b, a * a) + fma(dd, dd, cc * cc),
  "interleave (r2+y2)+(x2+z2)": (a * a + cc * cc) + (b * b + dd * dd),
  "interleave_fma fma(y,y,r2)+fma(z,z,x2)": fma(cc, cc, a * a) + fma(dd, dd, b * b),
  "fma_from_zero chain": fma(dd, dd, fma(cc, cc, fma(b, b, fma(a, a, np.zeros_like(a))))),
  "float64 sum": (a.astype(f64) ** 2 + b.astype(f64) ** 2 + cc.astype(f64) ** 2 + dd.astype(f64) ** 2).astype(f32),
}
for name, n2 in cands.items():
    n = np.maximum(np.sqrt(n2), f32(1e-12))
    out = g / n[:, None]
    print(f"  {name:45s} mismatching rows {int((out != t).any(axis=1).sum())}")
    outm = g * (f32(1.0) / n)[:, None]
    print(f"  {name:45s} (mul by reciprocal) {int((outm != t).any(axis=1).sum())}")
