"""Developer check on the GPU box: our CUDA path vs the committed golden vectors (which came from the reference)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "tests/golden", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import cases  # noqa: E402
import make_golden  # noqa: E402
import ours  # noqa: E402

names = sys.argv[1:] or list(cases.CASES)
for name in names:
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.isfile(path):
        print("no golden for", name)
        continue
    ref = dict(np.load(path))
    c, scene, cam, bg, dL, extra = cases.build_inputs(name)
    args, out, fwd = ours.run_forward(scene, cam, bg, extra)
    fwd["borderline"] = np.zeros((c["H"], c["W"]), bool)
    bwd = ours.run_backward(args, out, dL, c["lam"]) if c["backward"] else None
    make_golden.compare(name, ref, fwd, bwd)
