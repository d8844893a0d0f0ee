#!/usr/bin/env python
"""One forward + backward of BASELINE config C1 (10k Gaussians, 256x256) plus a small quantised / masked case through the
C ABI — the workload `compute-sanitizer` is pointed at (tools/gpu_sanitizer.sh).  Prints a checksum so a silent no-op is visible."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("reduced-3dgs_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from gs_b200 import synth  # noqa: E402
import ours  # noqa: E402


def main():
    W, H = synth.config_image("C1")
    scene = synth.config_scene("C1")
    cam = synth.make_camera(W, H)
    bg = torch.tensor([0.1, 0.2, 0.3])
    dL = synth.grad_image(W, H, 2)
    args, out, fwd = ours.run_forward(scene, cam, bg)
    g = ours.run_backward(args, out, dL, 0.05)
    print("C1 R", fwd["num_rendered"], "colour sum", float(fwd["color"].sum()), "grad sum", float(sum(abs(v).sum() for v in g.values())))
    # quantised + prune mask + non-multiple-of-16 image: the fused de-quantisation and the masked path under the sanitizer too
    W2, H2 = 200, 120
    s2 = synth.make_scene(4000, 3, mixed_degrees=True, box=(1.9 * W2 / H2, 1.9, 1.0), log_scale_mean=math.log(0.05))
    q = synth.quantise_scene(s2)
    deq = q.to("cuda").dequantise()
    s2 = synth.Scene(*[getattr(deq, f).cpu() for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees")])
    mask = synth.prune_mask(s2.P, 4)
    cam2 = synth.make_camera(W2, H2)
    a2, o2, f2 = ours.run_forward(s2, cam2, bg, prune_mask=mask, quant=q)
    g2 = ours.run_backward(a2, o2, synth.grad_image(W2, H2, 5), 0.0, prune_mask=mask, quant=q)
    print("quant+mask R", f2["num_rendered"], "colour sum", float(f2["color"].sum()), "grad sum", float(sum(abs(v).sum() for v in g2.values())))


if __name__ == "__main__":
    main()
