"""COMPARATOR (not product): NMR-style silhouette renderer = scalar CUDA kernels organised like the upstream
`neural_renderer` extension (baseline/nmr_style/nmr_style.cu) glued with eager PyTorch exactly as upstream does
(face doubling by torch.cat, vertices_to_faces gather, index flip, avg_pool2d). Used only by
scripts/bench_raster_vs_nmr_style.py and tests/test_nmr_style_gpu.py."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnmr_style.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "nmr_style.cu")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
                               "-shared", "-Xcompiler", "-fPIC", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        p, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        _lib.nmrs_forward_face_index_map.argtypes = [p, p, i, i, i, f, f, p, p, p, p]
        _lib.nmrs_forward_face_index_map_fast.argtypes = [p, p, i, i, i, f, f, p, p, p, p, p]
        _lib.nmrs_backward_pixel_map.argtypes = [p, p, p, p, i, i, i, f, p, p]
    return _lib


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, faces, image_size, near, far, eps, fast=False):
        f = faces.detach().contiguous().float()
        B, nf = f.shape[:2]
        dev = f.device
        face_index = torch.empty(B, image_size, image_size, dtype=torch.int32, device=dev)
        weight = torch.empty(B, image_size, image_size, 3, device=dev)
        depth = torch.empty(B, image_size, image_size, device=dev)
        inv = torch.empty(B, nf, 9, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        if fast:   # face-parallel over the pixel bounding box (the multiperson fork's organisation)
            keys = torch.empty(B, image_size, image_size, dtype=torch.int64, device=dev)
            rc = lib().nmrs_forward_face_index_map_fast(f.data_ptr(), inv.data_ptr(), B, nf, image_size, near, far,
                                                        keys.data_ptr(), face_index.data_ptr(), weight.data_ptr(),
                                                        depth.data_ptr(), s)
        else:      # pixel-parallel over all faces (the original extension)
            rc = lib().nmrs_forward_face_index_map(f.data_ptr(), inv.data_ptr(), B, nf, image_size, near, far,
                                                   face_index.data_ptr(), weight.data_ptr(), depth.data_ptr(), s)
        assert rc == 0
        alpha = (face_index >= 0).float()
        ctx.save_for_backward(f, face_index, alpha)
        ctx.meta = (image_size, eps)
        ctx.mark_non_differentiable(face_index)
        return alpha, face_index

    @staticmethod
    def backward(ctx, grad_alpha, _):
        f, face_index, alpha = ctx.saved_tensors
        image_size, eps = ctx.meta
        B, nf = f.shape[:2]
        g = grad_alpha.contiguous().float()
        grad_faces = torch.empty_like(f)
        rc = lib().nmrs_backward_pixel_map(f.data_ptr(), face_index.data_ptr(), alpha.data_ptr(), g.data_ptr(), B, nf,
                                           image_size, eps, grad_faces.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return grad_faces, None, None, None, None, None


def render_silhouettes(ndc, faces, image_size=256, anti_aliasing=True, near=0.1, far=100.0, eps=1e-4,
                       return_face_index=False, fast=False):
    """ndc [B,V,3] (already projected), faces [B,F,3] -> alpha [B,R,R], the upstream way."""
    faces = torch.cat((faces, faces[:, :, [2, 1, 0]]), dim=1).long()
    B, V = ndc.shape[:2]
    idx = faces + (torch.arange(B, device=ndc.device) * V)[:, None, None]
    faces_v = ndc.reshape(B * V, 3)[idx]
    S = image_size * 2 if anti_aliasing else image_size
    alpha, fi = _Rasterize.apply(faces_v, S, near, far, eps, fast)
    alpha = alpha.flip(1)
    if anti_aliasing:
        alpha = torch.nn.functional.avg_pool2d(alpha[:, None], 2)[:, 0]
    return (alpha, fi) if return_face_index else alpha
