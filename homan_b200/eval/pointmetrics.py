"""Metrics of the reference's evaluation (homan/eval/pointmetrics.py) on the SDF and nearest-point kernels.

`get_point_metrics(gt_points, pred_points)` (pointmetrics.py:17-45): symmetric chamfer distance (pytorch3d's
`chamfer_distance(..., batch_reduction=None)`: mean squared nearest-neighbour distance, both directions added), ADD-S
(mean nearest-neighbour distance from the ground truth to the estimate, bop_toolkit), mean vertex distance when the two
sets correspond. `get_align_metrics` (pointmetrics.py:62-99): the same after centring on the hand and rescaling the
prediction to the ground-truth hand size. Nearest neighbours by `hm_nearest_point`.

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


def nearest_dist2(a, b):
    """Squared distance from every point of a [B,N,3] to its nearest point of b [B,M,3] -> [B,N]."""
    if not (a.is_cuda and b.is_cuda):
        raise _lib.HomanB200Error("homan_b200.eval needs CUDA tensors (there is no CPU path)")
    a, b = a.detach().float().contiguous(), b.detach().float().contiguous()
    out = torch.empty(a.shape[0], a.shape[1], device=a.device)
    call("hm_nearest_point", ptr(a), ptr(b), a.shape[0], a.shape[1], b.shape[1], ptr(out), None, current_stream())
    return out


def chamfer_distance(x, y):
    """pytorch3d.loss.chamfer_distance(x, y, batch_reduction=None)[0] with its defaults (squared L2, point mean)."""
    return nearest_dist2(x, y).mean(1) + nearest_dist2(y, x).mean(1)


def get_point_metrics(gt_points, pred_points):
    """pointmetrics.py:17-45 -> {"chamfer_dists", "add-s", "verts_dists"}: lists of B floats."""
    gt, pred = gt_points.cuda(), pred_points.cuda()
    adis = nearest_dist2(gt, pred).sqrt().mean(1)
    results = {"chamfer_dists": chamfer_distance(gt, pred).cpu().numpy().tolist(), "add-s": adis.cpu().numpy().tolist()}
    if gt.shape[1] == pred.shape[1]:   # vertex assignments
        results["verts_dists"] = (gt.float() - pred.float()).norm(2, -1).mean(-1).cpu().numpy().tolist()
    else:
        results["verts_dists"] = list(results["add-s"])
    return results


def get_align_metrics(gt_hand_verts, pred_hand_verts, gt_obj_verts, pred_obj_verts):
    """pointmetrics.py:62-99 (one hand per scene): hand-centred, hand-scale-aligned vertex / chamfer errors. As upstream,
    the prediction is centred on the *ground-truth* hand centroid (pointmetrics.py:70: `pred_cent` is computed from
    `gt_hand_verts`)."""
    if gt_hand_verts.shape[0] != gt_obj_verts.shape[0]:
        raise NotImplementedError("homan_b200.eval.get_align_metrics: one hand per scene")
    gh, ph, go, po = (t.cuda().float() for t in (gt_hand_verts, pred_hand_verts, gt_obj_verts, pred_obj_verts))
    gt_cent = gh.mean(1, keepdim=True)
    pred_cent = gh.mean(1, keepdim=True)
    gh_c, go_c, ph_c, po_c = gh - gt_cent, go - gt_cent, ph - pred_cent, po - pred_cent
    gt_scale = torch.sqrt((gh_c.norm(2, -1) ** 2).sum(1) / gh.shape[1])
    pred_scale = torch.sqrt((ph_c.norm(2, -1) ** 2).sum(1) / ph.shape[1])
    ratio = (gt_scale / pred_scale).view(-1, 1, 1)
    ph_cs, po_cs = ph_c * ratio, po_c * ratio
    return {"hand_mean_aligned": (gh_c - ph_cs).norm(2, -1).mean(-1).cpu().numpy().tolist(),
            "obj_chamfer_aligned": chamfer_distance(po_cs, go_c).cpu().numpy().tolist()}
