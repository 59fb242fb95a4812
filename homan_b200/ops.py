"""torch.autograd.Function wrappers over the C ABI (libhoman_b200.so). GPU only, no fallback.

Each op names the reference interface it stands in for; the arithmetic lives in homan_b200/csrc/*.cu.
"""
import torch

from . import _lib
from ._lib import call, current_stream, ptr

NEAR, FAR, RASTER_EPS = 0.1, 100.0, 1e-4


def _check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.HomanB200Error("homan_b200 ops need CUDA tensors (there is no CPU path)")


def _f32(t):
    return t.detach().contiguous().float()


# ------------------------------------------------------------------------------------------ projection
class _Project(torch.autograd.Function):
    """nr.projection (neural_renderer; call sites /root/reference/homan/losses.py:34-41)."""

    @staticmethod
    def forward(ctx, verts, K, R, t, dist, orig_size, eps):
        _check_cuda(verts, K)
        v = _f32(verts)
        K = _f32(K).view(-1, 3, 3)
        B, V = v.shape[:2]
        R = None if R is None else _f32(R).view(-1)
        t = None if t is None else _f32(t).view(-1)
        dist = None if dist is None else _f32(dist).view(-1, 5)
        ndc = torch.empty_like(v)
        call("hm_project_fwd", ptr(v), ptr(K), K.shape[0], ptr(R), ptr(t), ptr(dist),
             0 if dist is None else dist.shape[0], float(orig_size), float(eps), B, V, ptr(ndc), current_stream())
        ctx.save_for_backward(v, K, R, t)
        ctx.meta = (float(orig_size), float(eps), dist is not None and bool((dist != 0).any()))
        return ndc

    @staticmethod
    def backward(ctx, grad_ndc):
        v, K, R, t = ctx.saved_tensors
        orig_size, eps, has_dist = ctx.meta
        if has_dist:
            raise _lib.HomanB200Error("projection backward supports zero distortion only")
        B, V = v.shape[:2]
        g = _f32(grad_ndc)
        out = torch.empty_like(v)
        call("hm_project_bwd", ptr(v), ptr(K), K.shape[0], ptr(R), ptr(t), orig_size, eps, B, V, ptr(g), ptr(out), 0,
             current_stream())
        return out, None, None, None, None, None, None


def project(verts, K, R=None, t=None, dist_coeffs=None, orig_size=1.0, eps=1e-9):
    return _Project.apply(verts, K, R, t, dist_coeffs, orig_size, eps)


# ------------------------------------------------------------------------------------------ rasteriser
class RasterBuffers:
    """Scratch of one silhouette render (records, bboxes, face_index, coverage / sweep bitmaps)."""

    def __init__(self, B, V, F, image_size, anti_aliasing, device):
        self.B, self.V, self.F, self.image_size, self.aa = B, V, F, image_size, bool(anti_aliasing)
        S = image_size * 2 if anti_aliasing else image_size
        self.S = S
        self.records = torch.empty(B * F * 224, dtype=torch.uint8, device=device)   # HM_FACE_RECORD_BYTES
        self.bboxes = torch.empty(((B * F * 8 + 15) // 16) * 16 + B * 16, dtype=torch.uint8, device=device)   # HM_FACE_BBOX_BUFFER_BYTES
        self.face_index = torch.empty(B, S, S, dtype=torch.int32, device=device)
        self.alpha = torch.empty(B, image_size, image_size, dtype=torch.float32, device=device)
        self.cov_row = torch.empty(B, S, S // 32, dtype=torch.int32, device=device)
        self.cov_col = torch.empty(B, S, S // 32, dtype=torch.int32, device=device)
        self.m_row = torch.empty(B, 2, S, S // 32, dtype=torch.int32, device=device)
        self.m_col = torch.empty(B, 2, S, S // 32, dtype=torch.int32, device=device)
        self.runs = torch.empty(B, 4, S, 8, 2, dtype=torch.int32, device=device)   # HM_RASTER_RUN_CAP = 8
        self.run_counts = torch.empty(B, 4, S, dtype=torch.int32, device=device)
        self.face_vis = torch.empty(B, (2 * F + 31) // 32, dtype=torch.int32, device=device)   # HM_FACE_VIS_WORDS(F)
        self.cov_blocks = torch.empty(B, S // 8, S // 32, dtype=torch.uint8, device=device)


def raster_forward(buf, ndc, faces, fill_back=True, near=NEAR, far=FAR):
    """ndc [B,V,3] fp32, faces [1|B,F,3] int32 -> buf.alpha [B,R,R] (+ face_index, coverage)."""
    s = current_stream()
    call("hm_raster_setup", ptr(ndc), ptr(faces), faces.shape[0], buf.B, buf.V, buf.F, buf.image_size, int(buf.aa),
         int(fill_back), ptr(buf.records), ptr(buf.bboxes), s)
    call("hm_raster_sil_fwd", ptr(buf.records), ptr(buf.bboxes), buf.B, buf.F, buf.image_size, int(buf.aa),
         float(near), float(far), ptr(buf.face_index), ptr(buf.alpha), ptr(buf.cov_row), ptr(buf.cov_col),
         ptr(buf.face_vis), ptr(buf.cov_blocks), s)
    return buf.alpha


def raster_backward(buf, grad_alpha, grad_ndc, eps=RASTER_EPS, grad_fixed=None, prepared=False):
    """grad_alpha [B,R,R] -> grad_ndc [B,V,3] += (approximate NMR gradient, x / y slots). `grad_fixed` (int64 [B,V,3],
    zeroed): order-independent fixed-point accumulation (test mode), folded into grad_ndc afterwards. `prepared`: the
    sweep masks / run lists of `buf` were already produced for this gradient (hm_sil_loss_prep)."""
    s = current_stream()
    if not prepared:
        call("hm_raster_grad_prep", ptr(grad_alpha), ptr(buf.cov_row), ptr(buf.cov_col), buf.B, buf.image_size,
             int(buf.aa), ptr(buf.m_row), ptr(buf.m_col), ptr(buf.runs), ptr(buf.run_counts), s)
    call("hm_raster_sil_bwd", ptr(buf.records), ptr(buf.bboxes), ptr(buf.face_index), ptr(grad_alpha),
         ptr(buf.cov_row), ptr(buf.cov_col), ptr(buf.face_vis), ptr(buf.cov_blocks), ptr(buf.m_row), ptr(buf.m_col),
         ptr(buf.runs), ptr(buf.run_counts),
         buf.B, buf.V, buf.F, buf.image_size,
         int(buf.aa), float(eps), ptr(grad_ndc), ptr(grad_fixed), s)
    if grad_fixed is not None:
        call("hm_fold_fixed", ptr(grad_fixed), grad_fixed.numel(), ptr(grad_ndc), s)
    return grad_ndc


class _RasterizeSilhouettes(torch.autograd.Function):
    """fill_back + vertices_to_faces + rasterize_silhouettes of neural_renderer
    (/root/reference/homan/losses.py:187 via Renderer.render_silhouettes)."""

    @staticmethod
    def forward(ctx, ndc, faces, image_size, anti_aliasing, fill_back, near, far, eps):
        _check_cuda(ndc, faces)
        ndc_c = _f32(ndc)
        faces_c = faces.detach().contiguous().int()
        if faces_c.dim() == 2:
            faces_c = faces_c[None]
        B, V = ndc_c.shape[:2]
        F = faces_c.shape[1]
        buf = RasterBuffers(B, V, F, image_size, anti_aliasing, ndc_c.device)
        raster_forward(buf, ndc_c, faces_c, fill_back, near, far)
        ctx.buf = buf
        ctx.eps = eps
        ctx.mark_non_differentiable(buf.face_index)
        return buf.alpha, buf.face_index

    @staticmethod
    def backward(ctx, grad_alpha, _grad_fi):
        buf = ctx.buf
        g = _f32(grad_alpha)
        grad_ndc = torch.zeros(buf.B, buf.V, 3, dtype=torch.float32, device=g.device)
        raster_backward(buf, g, grad_ndc, ctx.eps)
        return grad_ndc, None, None, None, None, None, None, None


def rasterize_silhouettes(ndc, faces, image_size=256, anti_aliasing=True, fill_back=True, near=NEAR, far=FAR,
                          eps=RASTER_EPS, return_face_index=False):
    alpha, face_index = _RasterizeSilhouettes.apply(ndc, faces, image_size, anti_aliasing, fill_back, near, far, eps)
    return (alpha, face_index) if return_face_index else alpha


# ------------------------------------------------------------------------------------------ RGB / depth (visualisation)
def lighting(faces_xyz, colours, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """nr.lighting for flat per-face colours: faces_xyz [B,nf,3,3] (3-D corners), colours [B,nf,3] -> lit colours."""
    dev = faces_xyz.device
    ca = torch.as_tensor(color_ambient, dtype=torch.float32, device=dev).view(1, 1, 3)
    cd = torch.as_tensor(color_directional, dtype=torch.float32, device=dev).view(1, 1, 3)
    d = torch.as_tensor(direction, dtype=torch.float32, device=dev).view(1, 1, 3)
    light = torch.zeros_like(colours)
    if intensity_ambient != 0:
        light = light + intensity_ambient * ca
    if intensity_directional != 0:
        n = torch.cross(faces_xyz[:, :, 0] - faces_xyz[:, :, 1], faces_xyz[:, :, 2] - faces_xyz[:, :, 1], dim=2)
        n = torch.nn.functional.normalize(n, dim=2, eps=1e-5)
        light = light + intensity_directional * cd * torch.relu((n * d).sum(2, keepdim=True))
    return colours * light


def face_lighting(verts, faces, colours, fill_back=True, intensity_ambient=0.5, intensity_directional=0.5,
                  color_ambient=(1, 1, 1), color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """nr.lighting for flat per-face colours, in the kernel (hm_face_lighting): verts [B,V,3], faces [1|B,F,3],
    colours [1|B,F,3] -> lit colours [B,2F,3] in the doubled numbering (f, F + f = reversed copy)."""
    import ctypes
    _check_cuda(verts, faces, colours)
    v = _f32(verts)
    f = faces.detach().contiguous().int()
    f = f[None] if f.dim() == 2 else f
    c = _f32(colours).view(-1, f.shape[1], 3)
    B, V, F = v.shape[0], v.shape[1], f.shape[1]
    lit = torch.empty(B, 2 * F, 3, device=v.device)
    vec = lambda x: (ctypes.c_float * 3)(*[float(y) for y in x])  # noqa: E731  (host vectors, read at enqueue time)
    ca, cd, d = vec(color_ambient), vec(color_directional), vec(direction)
    call("hm_face_lighting", ptr(v), ptr(f), f.shape[0], ptr(c), c.shape[0], B, V, F, int(bool(fill_back)),
         float(intensity_ambient), float(intensity_directional), ctypes.addressof(ca), ctypes.addressof(cd),
         ctypes.addressof(d), ptr(lit), current_stream())
    return lit


def render_rgbd(ndc, faces, lit_colours, image_size=256, anti_aliasing=True, fill_back=True, near=NEAR, far=FAR,
                background_color=(0, 0, 0)):
    """Forward of nr.rasterize_rgbad for texture_size 1 (visualisation, no gradient): ndc [B,V,3] projected
    vertices, faces [1|B,F,3], lit_colours [1|B,2F,3] (doubled numbering) -> rgb [B,3,R,R], depth [B,R,R],
    alpha [B,R,R]."""
    _check_cuda(ndc, faces, lit_colours)
    with torch.no_grad():
        ndc_c = _f32(ndc)
        faces_c = faces.detach().contiguous().int()
        if faces_c.dim() == 2:
            faces_c = faces_c[None]
        B, V = ndc_c.shape[:2]
        F = faces_c.shape[1]
        buf = RasterBuffers(B, V, F, image_size, anti_aliasing, ndc_c.device)
        raster_forward(buf, ndc_c, faces_c, fill_back, near, far)
        col = _f32(lit_colours).view(-1, 2 * F, 3)
        rgb = torch.empty(B, 3, image_size, image_size, device=ndc_c.device)
        depth = torch.empty(B, image_size, image_size, device=ndc_c.device)
        bg = [float(x) for x in background_color]
        call("hm_raster_shade", ptr(buf.records), ptr(buf.face_index), ptr(col), col.shape[0], B, F, image_size,
             int(bool(anti_aliasing)), float(far), bg[0], bg[1], bg[2], ptr(rgb), ptr(depth), current_stream())
        return rgb, depth, buf.alpha
