"""Drop-in for the `neural_renderer` package on the reference's path (silhouette mode):
`nr.renderer.Renderer(image_size, K, R, t, orig_size, ...)(vertices, faces, mode="silhouettes", K=...)`
(/root/reference/homan/losses.py:73-77,172-176,187; homan/homan.py:168-176) and `nr.projection`
(/root/reference/homan/losses.py:34-41), plus the forward of the RGB / depth render the reference uses for
visualisation (`mode=None`, `.render`; /root/reference/homan/homan.py:510-613) for texture_size-1 textures."""
import sys
import types

import numpy as np
import torch
from torch import nn

from .. import ops


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    return ops.project(vertices, K, R, t, dist_coeffs, orig_size, eps)


def vertices_to_faces(vertices, faces):
    bs, nv = vertices.shape[:2]
    faces = faces.long() + (torch.arange(bs, dtype=torch.long, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces]


class Renderer(nn.Module):
    def __init__(self, image_size=256, anti_aliasing=True, background_color=(0, 0, 0), fill_back=True,
                 camera_mode="projection", K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 perspective=True, viewing_angle=30, camera_direction=(0, 0, 1), near=0.1, far=100,
                 light_intensity_ambient=0.5, light_intensity_directional=0.5, light_color_ambient=(1, 1, 1),
                 light_color_directional=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        if camera_mode != "projection":
            raise NotImplementedError("homan_b200 Renderer: camera_mode='projection' only")
        self.image_size, self.anti_aliasing, self.fill_back = image_size, anti_aliasing, fill_back
        self.background_color, self.camera_mode = background_color, camera_mode
        cv = lambda x: torch.from_numpy(x).float().cuda() if isinstance(x, np.ndarray) else x  # noqa: E731
        self.K, self.R, self.t = cv(K), cv(R), cv(t)
        self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros(1, 5, device="cuda")
        self.orig_size, self.near, self.far = orig_size, near, far
        self.light_intensity_ambient, self.light_intensity_directional = light_intensity_ambient, light_intensity_directional
        self.light_color_ambient, self.light_color_directional = light_color_ambient, light_color_directional
        self.light_direction = light_direction
        self.rasterizer_eps = 1e-3

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None,
                orig_size=None):
        if mode == "silhouettes":
            return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)
        if mode is None:
            return self.render(vertices, faces, textures, K, R, t, dist_coeffs, orig_size)
        raise NotImplementedError("homan_b200 Renderer: mode=None (rgb, depth, alpha) or mode='silhouettes'")

    def render(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        """(rgb [B,3,R,R], depth [B,R,R], alpha [B,R,R]) for texture_size-1 textures [B,F,1,1,1,3]: flat lighting on the
        3-D faces, projection, z-buffer, colour / depth of the winning face. Forward only (visualisation)."""
        if textures.shape[2] != 1:
            raise NotImplementedError("homan_b200 Renderer.render: texture_size 1 (flat colour per face) only")
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        orig_size = self.orig_size if orig_size is None else orig_size
        with torch.no_grad():
            F = faces.shape[-2]
            f = faces if faces.dim() == 3 else faces[None]   # [1|B,F,3]
            colours = textures.reshape(textures.shape[0], F, 3).float()
            lit = ops.face_lighting(vertices, f, colours, self.fill_back, self.light_intensity_ambient,
                                    self.light_intensity_directional, self.light_color_ambient,
                                    self.light_color_directional, self.light_direction)
            ndc = ops.project(vertices, K, R, t, dist_coeffs, orig_size)
            return ops.render_rgbd(ndc, f, lit, self.image_size, self.anti_aliasing, self.fill_back, self.near,
                                   self.far, self.background_color)

    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        orig_size = self.orig_size if orig_size is None else orig_size
        ndc = ops.project(vertices, K, R, t, dist_coeffs, orig_size)
        return ops.rasterize_silhouettes(ndc, faces, self.image_size, self.anti_aliasing, self.fill_back)


renderer = types.ModuleType(__name__ + ".renderer")
renderer.Renderer = Renderer
sys.modules[renderer.__name__] = renderer
