"""Minimal input for `compute-sanitizer --tool racecheck`: one small silhouette forward / backward (shared-memory
z-buffer, hi-z summary, item queues, counting sort) - racecheck is ~100x slower than memcheck."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from homan_b200 import ops, synth  # noqa: E402

verts, faces = synth.make_object("ellipsoid80")
v = torch.from_numpy(verts.astype(np.float32))[None] * 6 + torch.tensor([0.0, 0.0, 1.0])
ndc = v.clone()
ndc[..., :2] = v[..., :2] / v[..., 2:] * 1.2
ndc = ndc.cuda().requires_grad_()
f = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
alpha = ops.rasterize_silhouettes(ndc, f, 64, True)
target = torch.roll(alpha.detach(), (3, -4), (1, 2)).round()
alpha.backward(2 * (alpha.detach() - target) / target.numel())
torch.cuda.synchronize()
print("race input ok", float(alpha.sum()), float(ndc.grad.abs().max()))
