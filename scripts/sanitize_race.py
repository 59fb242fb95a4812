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

# the fused loss + sweep-list kernel (shared-memory byte map, ballot masks) and the backward on its lists
from homan_b200._lib import call, current_stream, ptr  # noqa: E402
buf = ops.RasterBuffers(1, ndc.shape[1], f.shape[1], 64, True, "cuda")
ops.raster_forward(buf, ndc.detach(), f)
tgt = target.to(torch.int8).contiguous()
norm = torch.full((1,), 1e-4, device="cuda")
part = torch.zeros(1, 16, device="cuda")
ga = torch.empty(1, 64, 64, device="cuda")
call("hm_sil_loss_prep", ptr(buf.alpha), ptr(tgt), ptr(norm), 1.0, 1, 64, 1, ptr(part), 16, part.data_ptr() + 4, 16, ptr(ga),
     ptr(buf.cov_row), ptr(buf.cov_col), ptr(buf.m_row), ptr(buf.m_col), ptr(buf.runs), ptr(buf.run_counts), current_stream())
gn = torch.zeros(1, ndc.shape[1], 3, device="cuda")
ops.raster_backward(buf, ga, gn, prepared=True)
torch.cuda.synchronize()
print("fused prep ok", float(part[0, 0]), float(gn.abs().max()))
