"""ORACLE (test infrastructure; never imported by the product).

CPU stand-in for `from sdf import SDF` (hassony2/multiperson `sdf/`), used by the reference at
/root/reference/homan/interactions/scenesdf.py:9,32,119.  Semantics: SURVEY.md Appendix A.4
(third-party, parity unpinned).
"""
import torch
from torch import nn

from . import build as _build


class SDF(nn.Module):
    def forward(self, faces, vertices, grid_size=32):
        """faces [F,3] int32, vertices [B,V,3] float32 in [-1,1]^3 -> phi [B,G,G,G] (inside positive)."""
        lib = _build.lib()
        faces = faces.detach().contiguous().int()
        vertices = vertices.detach().contiguous().float()
        B, V = vertices.shape[:2]
        phi = torch.zeros(B, grid_size, grid_size, grid_size, dtype=torch.float32)
        lib.sdf_grid(faces.data_ptr(), faces.shape[0], vertices.data_ptr(), B, V, grid_size, phi.data_ptr())
        return phi
