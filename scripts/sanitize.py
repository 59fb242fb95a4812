"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): a few eager iterations of the
fused engine on the tiny workload plus one raster forward/backward through the autograd ops."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from homan_b200 import synth
from homan_b200.engine import FitEngine
from homan_b200.workload import make_workload

asset = synth.make_mano_asset(0, "right")
batch, lw = make_workload("tiny", mano_asset=asset)
eng = FitEngine(batch, lw, mano_asset=asset, use_graph=False)
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
print("total", eng.total.cpu().tolist())

# ---- rows added after the hot path: object-pose initialiser (off-screen / best-candidate kernels), RGB / depth
#      shade kernel, and the degenerate-face paths of the rasteriser (both windings front-facing, bit-line walk)
import numpy as np  # noqa: E402

from homan_b200 import ops, pose_optimization as po  # noqa: E402

verts, faces = synth.make_object("ellipsoid80")
K = np.array([[2.0, 0, 0.5], [0, 2.0, 0.5], [0, 0, 1]], np.float32)
rot = po.matrix_to_rot6d(po.compute_random_rotations(8, device="cuda"))
trans = torch.tensor([[0.0, 0.0, 0.5]], device="cuda").repeat(8, 1)
trans[3, 0] += 0.2
mask = np.zeros((256, 256), np.float32)
mask[96:160, 96:160] = 1
mask[:, 120:126] = -1
eng = po.PoseFitEngine(verts, faces, mask, K, rot, trans, use_graph=False)
for _ in range(2):
    eng.step()
eng.evaluate()
torch.cuda.synchronize()
print("pose total", eng.total.cpu().tolist()[:3], float(eng.best[0]))

rng = np.random.default_rng(0)
xy = rng.uniform(-0.8, 0.8, size=(40, 3, 2)).astype(np.float32)
xy[:10, 1] = xy[:10, 0]                                    # repeated vertex: both windings front-facing
xy[10:20, 2] = xy[10:20, 0] + 2 * (xy[10:20, 1] - xy[10:20, 0])  # collinear
z = rng.uniform(0.5, 1.5, size=(40, 3, 1)).astype(np.float32)
ndc = torch.from_numpy(np.concatenate((xy, z), 2).reshape(1, -1, 3)).cuda().requires_grad_()
f = torch.arange(120, dtype=torch.int32, device="cuda").view(1, 40, 3)
alpha = ops.rasterize_silhouettes(ndc, f, 64, True)
g = torch.rand_like(alpha) - 0.5                           # per-pixel random gradient: run lists overflow
alpha.backward(g)
lit = torch.rand(1, 80, 3, device="cuda")
rgb, depth, a2 = ops.render_rgbd(ndc.detach(), f, lit, 64, True)
torch.cuda.synchronize()
print("degenerate / shade ok", float(alpha.sum()), float(rgb.mean()), float(ndc.grad.abs().max()))
