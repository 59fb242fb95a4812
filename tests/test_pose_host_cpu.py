"""CPU: host-side helpers of the object-pose initialiser (homan_b200/pose_optimization.py) - the pieces of
find_optimal_pose that run before the kernels - against the reference's own functions when /root/reference is present
(build container) and against analytic properties everywhere."""
import math
import os

import numpy as np
import pytest
import torch

from homan_b200 import pose_optimization as po

REF = "/root/reference"


def test_random_rotations_are_rotations_and_cover_so3():
    g = torch.Generator().manual_seed(0)
    R = po.compute_random_rotations(4096, generator=g, device="cpu")
    assert R.shape == (4096, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand_as(R), atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(4096), atol=1e-5)
    # uniform on SO(3): the rotated z axis is uniform on the sphere (mean ~ 0, E[z^2] = 1/3)
    z = R[:, :, 2]
    assert z.mean(0).abs().max() < 0.05 and abs(float((z ** 2).mean()) - 1 / 3) < 0.02
    with pytest.raises(NotImplementedError):
        po.compute_random_rotations(2, upright=True, device="cpu")


def test_rot6d_roundtrip_and_orthonormalisation():
    g = torch.Generator().manual_seed(1)
    R = po.compute_random_rotations(64, generator=g, device="cpu")
    assert torch.allclose(po.rot6d_to_matrix(po.matrix_to_rot6d(R)), R, atol=1e-5)
    noisy = po.matrix_to_rot6d(R) * 1.7 + 0.05 * torch.randn(64, 3, 2, generator=g)
    Q = po.rot6d_to_matrix(noisy)
    assert torch.allclose(Q.transpose(1, 2) @ Q, torch.eye(3).expand_as(Q), atol=1e-5)


def test_tco_init_makes_the_projected_box_match_the_target():
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(5, 200, 3, generator=g) * 0.05
    K = torch.tensor([[600.0, 0, 320], [0, 600.0, 320], [0, 0, 1]])
    box = np.array([250.0, 200.0, 90.0, 120.0])  # xywh
    t = po.TCO_init_from_boxes_zup_autodepth(box, pts, K)
    proj = po.batch_proj2d(pts + t[:, None], K[None].repeat(5, 1, 1))
    lo, hi = proj.min(1)[0], proj.max(1)[0]
    centre = (lo + hi) / 2
    assert torch.allclose(centre, torch.tensor([295.0, 260.0]).expand_as(centre), atol=1.0)
    diag = (hi - lo).norm(dim=-1)
    assert torch.allclose(diag, torch.full((5,), math.hypot(90, 120)), rtol=0.02)


def test_k_crop_resize_maps_the_crop_onto_the_render_target():
    K = torch.tensor([[[600.0, 0, 320], [0, 600.0, 320], [0, 0, 1]]])
    x, y, b = 100.0, 120.0, 200.0
    Kc = po.get_K_crop_resize(K, torch.tensor([[x, y, x + b, y + b]]), [256])
    s = 256 / b
    assert np.isclose(Kc[0, 0, 0].item(), 600 * s) and np.isclose(Kc[0, 1, 1].item(), 600 * s)
    # a 3-D point projecting to pixel u in the full image lands on (u - x) * s (+ the upstream half-pixel terms)
    P = torch.tensor([0.02, -0.03, 0.5])
    u = (K[0] @ P)[:2] / P[2]
    uc = (Kc[0] @ P)[:2] / P[2]
    assert torch.allclose(uc, (u - torch.tensor([x, y])) * s, atol=1.0)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "homan")), reason="needs /root/reference (build container)")
def test_helpers_match_the_reference_functions(mano_assets, tmp_path):
    from oracle import refshim
    refshim.install(str(tmp_path), mano_assets)
    from homan.lib3d.optitrans import TCO_init_from_boxes_zup_autodepth as ref_tco
    from homan.utils import geometry as ref_geo
    g = torch.Generator().manual_seed(3)
    pts = torch.randn(7, 50, 3, generator=g) * 0.04
    K = np.array([[600.0, 0, 320], [0, 600.0, 320], [0, 0, 1]], dtype=np.float32)
    box = np.array([200.0, 260.0, 80.0, 60.0], dtype=np.float32)
    assert torch.allclose(po.TCO_init_from_boxes_zup_autodepth(box, pts, torch.from_numpy(K)[None]),
                          ref_tco(box, pts, torch.from_numpy(K)[None]), atol=1e-6)
    r6 = torch.randn(9, 3, 2, generator=g)
    assert torch.allclose(po.rot6d_to_matrix(r6), ref_geo.rot6d_to_matrix(r6), atol=1e-6)
    torch.manual_seed(5)
    a = ref_geo.compute_random_rotations(16)
    torch.manual_seed(5)
    b = po.compute_random_rotations(16, device="cpu")
    assert torch.allclose(a, b, atol=1e-6)
