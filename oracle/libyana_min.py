"""ORACLE (test infrastructure; never imported by the product).

The handful of `libyana` (github.com/hassony2/libyana, un-vendored) helpers that sit on the
hot path of the reference; semantics per SURVEY.md Appendix A.6 (parity unpinned).
Call sites: /root/reference/homan/losses.py:13-15,147,192,220,227;
/root/reference/homan/jointopt.py:15-16,52-53; /root/reference/homan/homan.py:21-23,160.
"""
import numpy as np
import torch


def tensorify(array, device=None):
    if isinstance(array, torch.Tensor):
        out = array
    else:
        out = torch.from_numpy(np.asarray(array))
        if out.dtype == torch.float64:
            out = out.float()
    if device is not None:
        out = out.to(device)
    return out


def numpify(tensor):
    if isinstance(tensor, torch.Tensor):
        return tensor.detach().cpu().numpy()
    return np.asarray(tensor)


def batch_proj2d(verts, camintr, camextr=None):
    """verts [B,V,3], camintr [B,3,3] -> [B,V,2] = (K v)_{xy} / (K v)_z."""
    if camextr is not None:
        raise NotImplementedError
    hom = camintr.bmm(verts.transpose(1, 2)).transpose(1, 2)
    return hom[:, :, :2] / hom[:, :, 2:]


def batch_mask_iou(ref, pred, eps=0.000001):
    """Per-image IoU of soft masks [B,H,W] -> [B]."""
    ref = ref.float()
    pred = pred.float()
    if ref.max() > 1 or ref.min() < 0:
        raise ValueError("ref not in [0,1]")
    if pred.max() > 1 or pred.min() < 0:
        raise ValueError("pred not in [0,1]")
    inter = (ref * pred).sum((1, 2))
    union = (ref + pred).clamp(0, 1).sum((1, 2))
    return inter / (union + eps)


def batch_pairwise_dist(x, y, use_cuda=False):
    """Squared distances [B,Nx,Ny] via rx + ry - 2 x.y (same formula as contactloss.py:60-79)."""
    xx = torch.bmm(x, x.transpose(2, 1))
    yy = torch.bmm(y, y.transpose(2, 1))
    zz = torch.bmm(x, y.transpose(2, 1))
    ix = torch.arange(0, x.shape[1], device=x.device)
    iy = torch.arange(0, y.shape[1], device=x.device)
    rx = xx[:, ix, ix].unsqueeze(1).expand_as(zz.transpose(2, 1))
    ry = yy[:, iy, iy].unsqueeze(1).expand_as(zz)
    return rx.transpose(2, 1) + ry - 2 * zz


def check_shape(tensor, exp_shape, name="tensor"):
    shape = tuple(tensor.shape)
    if len(shape) != len(exp_shape) or any(e != -1 and s != e for s, e in zip(shape, exp_shape)):
        raise ValueError(f"{name}: expected shape {exp_shape}, got {shape}")


def get_K_crop_resize(K, boxes, crop_resize, invert_xy=False):
    """libyana.lib3d.kcrop.get_K_crop_resize (call site /root/reference/homan/pose_optimization.py:247-249):
    intrinsics of the crop `boxes` (xyxy) resized to crop_resize; skew not handled. Recalled from upstream
    (cosypose-derived), parity unpinned. K [B,3,3], boxes [B,4] -> [B,3,3]."""
    assert K.shape[1:] == (3, 3) and boxes.shape[1:] == (4,)
    K, boxes = K.float(), boxes.float()
    new_K = K.clone()
    final_width, final_height = float(max(crop_resize)), float(min(crop_resize))
    crop_width, crop_height = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    crop_cj, crop_ci = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_width - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_height - 1) / 2 - crop_ci
    center_x, center_y = (crop_width - 1) / 2, (crop_height - 1) / 2
    scale_x, scale_y = final_width / crop_width, final_height / crop_height
    new_K[:, 0, 0] = scale_x * K[:, 0, 0]
    new_K[:, 1, 1] = scale_y * K[:, 1, 1]
    new_K[:, 0, 2] = (final_width - 1) / 2 + scale_x * (cx - center_x)
    new_K[:, 1, 2] = (final_height - 1) / 2 + scale_y * (cy - center_y)
    return new_K
