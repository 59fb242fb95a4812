"""Interaction metrics of the reference's evaluation (homan/eval/pointmetrics.py:102-124) on the SDF kernel.

`get_inter_metrics(verts_person, verts_object, faces_person, faces_object)` keeps the reference signature and return
dict: per scene the penetration depth of the hand into the object - the maximum over the hand vertices of the object's
clamped signed distance field (32^3 grid over the object's enlarged bounding cube, trilinear samples, metres), i.e.
`dist_values[(1, 0)].max(1)` of SDFSceneLoss - and whether there is contact at all (depth > 0). One `hm_sdf_pair`
launch (the sparse kernel of the collision loss); GPU only."""
import torch

from .. import _lib
from .._lib import call, current_stream, ptr

SDF_GRID, SDF_SCALE_FACTOR = 32, 0.2   # homan/interactions/scenesdf.py:14,77


def sdf_dist_values(verts_grid, faces_grid, verts_sampled):
    """dist_values[(grid mesh, sampled mesh)] of SDFSceneLoss.forward (scenesdf.py:128-146): [B, Vs], scene units."""
    if not (verts_grid.is_cuda and verts_sampled.is_cuda):
        raise _lib.HomanB200Error("homan_b200.eval needs CUDA tensors (there is no CPU path)")
    vg = verts_grid.detach().float().contiguous()
    vs = verts_sampled.detach().float().contiguous()
    fg = faces_grid.detach().to(vg.device).int().contiguous().view(-1, 3)
    B, Vg, Vs = vg.shape[0], vg.shape[1], vs.shape[1]
    phi = torch.empty(B, SDF_GRID ** 3, device=vg.device)
    partials = torch.zeros(B, 16, device=vg.device)
    out = torch.empty(B, Vs, device=vg.device)
    call("hm_sdf_pair", ptr(vg), ptr(fg), 1, ptr(vs), B, Vg, fg.shape[0], Vs, SDF_GRID, SDF_SCALE_FACTOR, 0.0,
         ptr(phi), ptr(partials), None, ptr(out), current_stream())
    return out


def get_inter_metrics(verts_person, verts_object, faces_person, faces_object):
    """verts_person [B*H,778,3] (H hands per scene, H <= 2), verts_object [B,Vo,3], faces_* [1|H,F,3].
    Returns {"pen_depths": [B floats], "has_contact": [B bools]} (pointmetrics.py:102-124)."""
    hand_nb = verts_person.shape[0] // verts_object.shape[0]
    if hand_nb == 2:   # both hands as one vertex set (the faces of the hands are not used by the metric's (1, 0) pair)
        verts_person = verts_person.view(verts_object.shape[0], -1, 3)
    elif hand_nb > 3 or hand_nb < 1:
        raise ValueError(f"Invalid hand nb {hand_nb}")
    depths = sdf_dist_values(verts_object, faces_object[0], verts_person).max(1)[0]
    return {"pen_depths": depths.cpu().numpy().tolist(), "has_contact": (depths > 0).cpu().numpy().tolist()}
