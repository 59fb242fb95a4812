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
