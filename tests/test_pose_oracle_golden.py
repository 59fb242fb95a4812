"""CPU: the oracle restatement of the object-pose multi-init fitter (oracle/pose_ref.py) against the golden vectors
recorded from the UNMODIFIED reference /root/reference/homan/pose_optimization.py (scripts/make_golden_pose.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_pose_init.npz")


def _golden():
    return np.load(GOLDEN)


def test_oracle_fit_follows_the_reference_trajectory():
    from oracle import pose_ref
    g = _golden()
    iters = int(g["iters"])
    out = pose_ref.fit(g["in_vertices"], g["in_faces"], g["in_mask"], g["K_roi"],
                       g["in_rotations_init"][:, :, :2], g["it_trans"][0], iters)
    ref_total = g["it_mask"] + g["it_offscreen"] + g["it_chamfer"]
    assert np.allclose(out["total"], ref_total, rtol=1e-5)          # same CPU arithmetic: equal to rounding
    assert np.allclose(out["iou"], g["it_iou"], rtol=1e-5)
    assert np.allclose(out["rotations"], g["plain_rotations"], atol=1e-6)
    assert np.allclose(out["translations"], g["plain_translations"], atol=1e-6)
    # sort_best=True: best-ever candidate first, then the candidates sorted by their last evaluated loss
    assert np.allclose(out["best"]["rot"], g["sorted_rotations"][0], atol=1e-6)
    order = np.argsort(ref_total[-1], kind="stable")
    assert np.allclose(out["rotations"][order][:-1], g["sorted_rotations"][1:], atol=1e-6)


def test_oracle_offscreen_term_and_gradients():
    from oracle import pose_ref
    g = _golden()
    rot = torch.from_numpy(g["po_rot6d"]).requires_grad_()
    tr = torch.from_numpy(g["po_trans"]).requires_grad_()
    ld, iou, _ = pose_ref.forward(torch.from_numpy(g["in_vertices"]), torch.from_numpy(g["in_faces"]).int(), g["in_mask"],
                                  torch.from_numpy(g["K_roi"]), rot, tr)
    assert np.allclose(ld["mask"].detach().numpy(), g["po_mask"])
    assert np.allclose(ld["offscreen"].detach().numpy(), g["po_offscreen"], rtol=1e-6)
    assert g["po_offscreen"][0] == 0 and (g["po_offscreen"][1:] > 0).all()
    assert np.allclose(iou.numpy(), g["po_iou"], rtol=1e-6)
    (ld["mask"] + ld["offscreen"]).sum().backward()
    assert np.allclose(rot.grad.numpy(), g["po_grad_rot"], rtol=1e-4, atol=1e-3)
    assert np.allclose(tr.grad.numpy(), g["po_grad_trans"], rtol=1e-4, atol=1e-3)


def test_kcrop_matches_the_crop_used_by_the_workload_generator():
    """get_K_crop_resize (recalled libyana semantics) maps the crop corner to -0.5 px and scales the focal."""
    from oracle import libyana_min
    K = torch.tensor([[[600.0, 0, 320], [0, 600.0, 320], [0, 0, 1]]])
    out = libyana_min.get_K_crop_resize(K, torch.tensor([[100.0, 120.0, 300.0, 320.0]]), [256])
    assert np.isclose(out[0, 0, 0].item(), 600 * 256 / 200) and np.isclose(out[0, 1, 1].item(), 600 * 256 / 200)
    # a point projecting to the crop's top-left pixel centre (100, 120) lands on pixel (0, 0) of the resized crop
    # up to the half-pixel convention of the upstream formula
    u = out[0, 0, 0].item() * (100 - 320) / 600 + out[0, 0, 2].item()
    assert abs(u) <= 1.0
