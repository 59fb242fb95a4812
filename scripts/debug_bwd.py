"""Development tool (GPU): the integer-coordinate stress scene through the product backward, the oracle and the
scalar model (per pass), worst vertices listed."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from homan_b200 import ops  # noqa: E402
from oracle import build as ob, nmr  # noqa: E402

proto = ctypes.CDLL(os.path.join(ROOT, "scripts", "proto", "libbwdproto.so"))
P = ctypes.c_void_p
proto.proto_pixel_map_bwd.argtypes = [P, P, P, P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                      ctypes.c_float, P, P]

rng = np.random.default_rng(2)
S = 256
n = 200
pix = rng.integers(8, S - 8, size=(n, 3, 2)).astype(np.float64)
pix[1::2] += 0.5
xy = (2 * pix + 1 - S) / S
z = rng.uniform(0.5, 1.5, size=(n, 3, 1))
verts = np.concatenate((xy, z), 2).reshape(-1, 3).astype(np.float32)
faces = np.arange(3 * n).reshape(n, 3)
ndc = torch.from_numpy(verts[None]).contiguous()
fb = torch.from_numpy(faces).long()[None]
f2 = torch.cat((fb, fb[:, :, [2, 1, 0]]), dim=1)
fv = nmr.vertices_to_faces(ndc, f2).contiguous().float()
lib = ob.lib()
fi = torch.empty(1, S, S, dtype=torch.int32)
lib.nmr_face_index_map(fv.data_ptr(), 1, 2 * n, S, 0.1, 100.0, fi.data_ptr(), None)
alpha = (fi >= 0).float()
a_img = torch.flip(alpha, dims=(1,))
target = torch.roll(a_img, shifts=(5, -7), dims=(1, 2)).round()
g_alpha = 2 * (a_img - target) / target[0].numel()
g = torch.flip(g_alpha, dims=(1,)).contiguous()


def to_verts(gf):
    out = torch.zeros(3 * n, 3)
    out.index_add_(0, f2[0].reshape(-1), gf[0].reshape(-1, 3))
    return out


ref = torch.zeros_like(fv)
lib.nmr_pixel_map_bwd(fv.data_ptr(), fi.data_ptr(), alpha.data_ptr(), g.data_ptr(), 1, 2 * n, S, 1e-4, ref.data_ptr())
parts = {}
for mode in (1, 2, 4):
    proto.proto_set_mode(mode)
    out = torch.zeros_like(fv)
    st = np.zeros(16, dtype=np.int64)
    proto.proto_pixel_map_bwd(fv.data_ptr(), fi.data_ptr(), alpha.data_ptr(), g.data_ptr(), 1, 2 * n, S, 1e-4, 128.0,
                              out.data_ptr(), st.ctypes.data)
    parts[mode] = to_verts(out)
ref_v = to_verts(ref)
nd = ndc.cuda().requires_grad_()
al, fi_d = ops.rasterize_silhouettes(nd, fb.int().cuda().contiguous(), S, False, return_face_index=True)
assert torch.equal(fi_d.cpu(), fi)
al.backward(g_alpha.cuda())
got = nd.grad[0].cpu()
role = os.environ.get("ROLE", "")
if role == "A":
    ref_v = parts[1]
elif role == "B":
    ref_v = parts[2] + parts[4]
np.savez(os.path.join(ROOT, "gpurun_out", f"debug_bwd_{role or 'all'}.npz"), got=got.numpy(), ref=ref_v.numpy(),
         O=parts[1].numpy(), S=parts[2].numpy(), I=parts[4].numpy(), fi=fi.numpy(), verts=verts, g=g.numpy())
err = (got - ref_v).abs()
print("max err", float(err.max()), "scale", float(ref_v.abs().max()))
idx = torch.argsort(err.reshape(-1), descending=True)[:12]
for i in idx:
    v, c = int(i) // 3, int(i) % 3
    print(f"vertex {v} (face {v // 3} corner {v % 3} parity {(v // 3) % 2}) comp {c}: got {got[v, c]:+.6e} ref {ref_v[v, c]:+.6e} "
          f"diff {got[v, c] - ref_v[v, c]:+.3e} | O {parts[1][v, c]:+.6e} S {parts[2][v, c]:+.6e} I {parts[4][v, c]:+.6e}")
