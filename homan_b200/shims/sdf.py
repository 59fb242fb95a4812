"""Drop-in for `from sdf import SDF` (/root/reference/homan/interactions/scenesdf.py:9,32,119):
SDF()(faces int32 [F,3], vertices fp32 [B,V,3] in [-1,1]^3, grid_size=32) -> phi [B,G,G,G], inside positive.
Dense evaluation (this is the drop-in; the fused engine evaluates the grid sparsely, see csrc/sdf.cu)."""
import torch
from torch import nn

from .._lib import call, current_stream, ptr


class SDF(nn.Module):
    def forward(self, faces, vertices, grid_size=32):
        faces = faces.detach().contiguous().int()
        vertices = vertices.detach().contiguous().float()
        B, V = vertices.shape[:2]
        phi = torch.empty(B, grid_size, grid_size, grid_size, dtype=torch.float32, device=vertices.device)
        call("hm_sdf_grid", ptr(faces), ptr(vertices), B, V, faces.shape[0], grid_size, ptr(phi), current_stream())
        return phi
