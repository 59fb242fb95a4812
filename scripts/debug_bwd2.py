"""Development: per-image difference between the raster backward and the NMR-style comparator on cfg3's object."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import nmr_style
from homan_b200 import ops, synth
from homan_b200.workload import CONFIGS
c = CONFIGS["cfg3"]
asset = synth.make_mano_asset(0, "right")
clip = synth.make_clip(c["T"], c["obj"], seed=c["seed"], mano_asset=asset)
inits = synth.make_inits(clip, c["P"], seed=c["seed"])
P, T = c["P"], c["T"]
R = np.einsum("vk,ptkj->ptvj", clip["obj_verts_can"], inits["obj_R"].astype(np.float32))
verts = (R + inits["obj_t"][:, :, None]).reshape(P * T, -1, 3).astype(np.float32)
faces, K = clip["obj_faces"], np.tile(clip["K_roi_obj"][None], (P, 1, 1, 1)).reshape(P * T, 3, 3)
v, Kd = torch.from_numpy(verts).cuda(), torch.from_numpy(K.astype(np.float32)).cuda()
f32 = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
ndc = ops.project(v, Kd, orig_size=1.0).detach()
B, V, F = ndc.shape[0], ndc.shape[1], faces.shape[0]
buf = ops.RasterBuffers(B, V, F, 256, True, ndc.device)
ops.raster_forward(buf, ndc, f32)
target = torch.roll(buf.alpha, shifts=(5, -7), dims=(1, 2)).round()
g = (2 * (buf.alpha - target) / (256 * 256 * T)).contiguous()
gn = torch.zeros(B, V, 3, device=ndc.device)
ops.raster_backward(buf, g, gn)
n = 64
nd = ndc[:n].clone().requires_grad_()
a = nmr_style.render_silhouettes(nd, f32.long().repeat(n, 1, 1), 256, True, fast=True)
(gc,) = torch.autograd.grad(a, nd, g[:n])
d = (gc - gn[:n]).abs().flatten(1).max(1)[0]
sc = gc.abs().flatten(1).max(1)[0]
print("per-image rel diff:", (d / sc).cpu().numpy().round(5))
w = int((d / sc).argmax())
dv = (gc[w] - gn[w]).abs()
idx = torch.nonzero(dv > 1e-3 * sc[w])
print("worst image", w, "vertices/components off:", idx[:20].cpu().numpy().tolist())
print("values cmp:", gc[w][idx[:8, 0], idx[:8, 1]].cpu().numpy(), "ours:", gn[w][idx[:8, 0], idx[:8, 1]].cpu().numpy())
# run counts: overflow lines?
rc = buf.run_counts[w].cpu().numpy()
print("lines with overflow (cnt==15):", int(((rc & 15) == 15).sum()), "max cnt", int((rc & 15).max()))
np.savez("gpurun_out/debug_bwd2.npz", ndc=ndc[w].cpu().numpy(), faces=faces, g=g[w].cpu().numpy(), gc=gc[w].cpu().numpy(), gn=gn[w].cpu().numpy())
