"""ORACLE (test infrastructure; never imported by the product).

CPU stand-in for the `neural_renderer` package (Kato NMR, PyTorch port of
hassony2/multiperson): the silhouette path and the forward of the RGB / depth render, as used by the reference at
/root/reference/homan/losses.py:34-41,73-77,172-176,187 and
/root/reference/homan/homan.py:168-176.  Semantics: SURVEY.md Appendix A.1-A.3
(third-party package absent from the reference tree: parity unpinned).
"""
import numpy as np
import torch
from torch import nn

from . import build as _build

DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100.0
DEFAULT_EPS = 1e-4


def _ptr(t):
    return t.data_ptr()


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """Pin-hole projection to NDC (Appendix A.2).  vertices [B,V,3], K [1|B,3,3], R [1,3,3], t [1,3]."""
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1 = dist_coeffs[:, None, 0]
    k2 = dist_coeffs[:, None, 1]
    p1 = dist_coeffs[:, None, 2]
    p2 = dist_coeffs[:, None, 3]
    k3 = dist_coeffs[:, None, 4]
    r = torch.sqrt(x_**2 + y_**2)
    x__ = x_ * (1 + k1 * (r**2) + k2 * (r**4) + k3 * (r**6)) + 2 * p1 * x_ * y_ + p2 * (r**2 + 2 * x_**2)
    y__ = y_ * (1 + k1 * (r**2) + k2 * (r**4) + k3 * (r**6)) + p1 * (r**2 + 2 * y_**2) + 2 * p2 * x_ * y_
    vertices = torch.stack([x__, y__, torch.ones_like(z)], dim=-1)
    vertices = torch.matmul(vertices, K.transpose(1, 2))
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.0) / orig_size
    v = 2 * (v - orig_size / 2.0) / orig_size
    return torch.stack([u, v, z], dim=-1)


def vertices_to_faces(vertices, faces):
    """[B,V,3], [B,F,3] int -> [B,F,3,3] (gather; backward = scatter-add)."""
    bs, nv = vertices.shape[:2]
    device = vertices.device
    faces = faces.long() + (torch.arange(bs, dtype=torch.long, device=device) * nv)[:, None, None]
    vertices = vertices.reshape((bs * nv, 3))
    return vertices[faces]


class _RasterizeSilhouette(torch.autograd.Function):
    """faces [B,nf,3,3] -> alpha [B,is,is] in the raster frame (row 0 = y -1, before the flip)."""

    @staticmethod
    def forward(ctx, faces, image_size, near, far, eps):
        lib = _build.lib()
        faces_c = faces.detach().contiguous().float()
        B, nf = faces_c.shape[:2]
        face_index = torch.empty(B, image_size, image_size, dtype=torch.int32)
        lib.nmr_face_index_map(_ptr(faces_c), B, nf, image_size, near, far, _ptr(face_index), None)
        alpha = (face_index >= 0).float()
        ctx.save_for_backward(faces_c, face_index, alpha)
        ctx.image_size = image_size
        ctx.eps = eps
        ctx.mark_non_differentiable(face_index)
        return alpha, face_index

    @staticmethod
    def backward(ctx, grad_alpha, _grad_fi):
        lib = _build.lib()
        faces_c, face_index, alpha = ctx.saved_tensors
        B, nf = faces_c.shape[:2]
        grad_alpha = grad_alpha.contiguous().float()
        grad_faces = torch.zeros_like(faces_c)
        lib.nmr_pixel_map_bwd(_ptr(faces_c), _ptr(face_index), _ptr(alpha), _ptr(grad_alpha), B, nf,
                              ctx.image_size, ctx.eps, _ptr(grad_faces))
        return grad_faces, None, None, None, None


def rasterize_silhouettes(faces, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                          eps=DEFAULT_EPS, return_face_index=False):
    """Appendix A.3: rasterise at 2x when anti-aliasing, vertical flip, 2x2 average pool."""
    is_ = image_size * 2 if anti_aliasing else image_size
    alpha, face_index = _RasterizeSilhouette.apply(faces, is_, near, far, eps)
    alpha = alpha[:, list(reversed(range(alpha.shape[1]))), :]
    if anti_aliasing:
        alpha = torch.nn.functional.avg_pool2d(alpha[:, None, :, :], kernel_size=(2, 2))[:, 0]
    if return_face_index:
        return alpha, face_index
    return alpha


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """nr.lighting (Appendix A.1): flat shading, light = ambient + directional * relu(n . direction) with the face normal
    n = normalize(cross(v0 - v1, v2 - v1), eps=1e-5). faces [B,nf,3,3] (3-D, before projection), textures
    [B,nf,ts,ts,ts,3] -> lit textures."""
    bs, nf = faces.shape[:2]
    ca = torch.as_tensor(color_ambient, dtype=torch.float32).view(1, 3)
    cd = torch.as_tensor(color_directional, dtype=torch.float32).view(1, 3)
    d = torch.as_tensor(direction, dtype=torch.float32).view(1, 1, 3)
    light = torch.zeros(bs, nf, 3)
    if intensity_ambient != 0:
        light = light + intensity_ambient * ca[:, None, :]
    if intensity_directional != 0:
        f = faces.reshape(bs * nf, 3, 3)
        normals = torch.nn.functional.normalize(torch.cross(f[:, 0] - f[:, 1], f[:, 2] - f[:, 1], dim=1), eps=1e-5)
        cos = torch.relu(torch.sum(normals.reshape(bs, nf, 3) * d, dim=2))
        light = light + intensity_directional * (cd[:, None, :] * cos[:, :, None])
    return textures * light[:, :, None, None, None, :]


def rasterize_rgbad(faces, textures, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                    background_color=(0, 0, 0)):
    """Forward of nr.rasterize_rgbad for texture_size 1 (flat colour per face; the reference's textures are
    [B,F,1,1,1,3], /root/reference/homan/homan.py:510-518): rgb [B,3,R,R], depth [B,R,R], alpha [B,R,R] after the
    vertical flip and, with anti-aliasing, the 2x2 average pool. Visualisation only: no backward."""
    if textures.shape[2] != 1:
        raise NotImplementedError("oracle rasterize_rgbad: texture_size 1 only")
    lib = _build.lib()
    is_ = image_size * 2 if anti_aliasing else image_size
    faces_c = faces.detach().contiguous().float()
    B, nf = faces_c.shape[:2]
    face_index = torch.empty(B, is_, is_, dtype=torch.int32)
    depth = torch.empty(B, is_, is_)
    lib.nmr_face_index_map(_ptr(faces_c), B, nf, is_, near, far, _ptr(face_index), _ptr(depth))
    covered = face_index >= 0
    colours = textures.detach().reshape(B, nf, 3).float()
    idx = face_index.clamp(min=0).long().view(B, -1)
    rgb = torch.gather(colours, 1, idx[:, :, None].expand(-1, -1, 3)).view(B, is_, is_, 3)
    bg = torch.as_tensor(background_color, dtype=torch.float32)
    rgb = torch.where(covered[..., None], rgb, bg.view(1, 1, 1, 3).expand_as(rgb))
    alpha = covered.float()
    flip = list(reversed(range(is_)))
    rgb = rgb.permute(0, 3, 1, 2)[:, :, flip, :]
    alpha, depth = alpha[:, flip, :], depth[:, flip, :]
    if anti_aliasing:
        pool = lambda x: torch.nn.functional.avg_pool2d(x, kernel_size=(2, 2))  # noqa: E731
        rgb, alpha, depth = pool(rgb), pool(alpha[:, None])[:, 0], pool(depth[:, None])[:, 0]
    return rgb, depth, alpha


class Renderer(nn.Module):
    """Subset of nr.renderer.Renderer used by the reference (Appendix A.1)."""

    def __init__(self, image_size=256, anti_aliasing=True, background_color=(0, 0, 0), fill_back=True,
                 camera_mode="projection", K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 near=0.1, far=100, light_intensity_ambient=0.5, light_intensity_directional=0.5,
                 light_color_ambient=(1, 1, 1), light_color_directional=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        self.background_color = background_color
        self.fill_back = fill_back
        self.camera_mode = camera_mode
        if camera_mode != "projection":
            raise ValueError("oracle Renderer supports camera_mode='projection' only")
        self.K, self.R, self.t = K, R, t
        if isinstance(self.K, np.ndarray):
            self.K = torch.from_numpy(self.K).float()
        if isinstance(self.R, np.ndarray):
            self.R = torch.from_numpy(self.R).float()
        if isinstance(self.t, np.ndarray):
            self.t = torch.from_numpy(self.t).float()
        self.dist_coeffs = dist_coeffs
        if dist_coeffs is None:
            self.dist_coeffs = torch.zeros(1, 5)
        self.orig_size = orig_size
        self.near = near
        self.far = far
        self.light_intensity_ambient = light_intensity_ambient
        self.light_intensity_directional = light_intensity_directional
        self.light_color_ambient = light_color_ambient
        self.light_color_directional = light_color_directional
        self.light_direction = light_direction
        self.rasterizer_eps = 1e-3

    def forward(self, vertices, faces, textures=None, mode=None, K=None, R=None, t=None, dist_coeffs=None,
                orig_size=None):
        if mode == "silhouettes":
            return self.render_silhouettes(vertices, faces, K, R, t, dist_coeffs, orig_size)
        if mode is None:
            return self.render(vertices, faces, textures, K, R, t, dist_coeffs, orig_size)
        raise NotImplementedError("oracle Renderer implements mode=None and mode='silhouettes'")

    def project_faces(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        if self.fill_back:
            faces = torch.cat((faces, faces[:, :, list(reversed(range(faces.shape[-1])))]), dim=1)
        K = self.K if K is None else K
        R = self.R if R is None else R
        t = self.t if t is None else t
        dist_coeffs = self.dist_coeffs if dist_coeffs is None else dist_coeffs
        orig_size = self.orig_size if orig_size is None else orig_size
        vertices = projection(vertices, K, R, t, dist_coeffs.to(vertices.device), orig_size)
        return vertices_to_faces(vertices, faces)

    def render_silhouettes(self, vertices, faces, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        faces_v = self.project_faces(vertices, faces, K, R, t, dist_coeffs, orig_size)
        return rasterize_silhouettes(faces_v, self.image_size, self.anti_aliasing)

    def render(self, vertices, faces, textures, K=None, R=None, t=None, dist_coeffs=None, orig_size=None):
        """Renderer.render (Appendix A.1): fill_back, flat lighting on the 3-D faces, projection, rasterize_rgbad ->
        (rgb [B,3,R,R], depth [B,R,R], alpha [B,R,R]). Used by the reference for visualisation
        (/root/reference/homan/homan.py:535, homan/utils/nmr_renderer.py:71,164,209)."""
        faces_l = faces
        if self.fill_back:
            faces_l = torch.cat((faces, faces[:, :, list(reversed(range(faces.shape[-1])))]), dim=1)
            textures = torch.cat((textures, textures.permute((0, 1, 4, 3, 2, 5))), dim=1)
        textures = lighting(vertices_to_faces(vertices, faces_l), textures, self.light_intensity_ambient,
                            self.light_intensity_directional, self.light_color_ambient, self.light_color_directional,
                            self.light_direction)
        faces_v = self.project_faces(vertices, faces, K, R, t, dist_coeffs, orig_size)
        return rasterize_rgbad(faces_v, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                               self.background_color)
