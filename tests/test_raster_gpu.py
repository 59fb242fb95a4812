"""GPU parity of the silhouette rasteriser (forward + approximate backward) and the projection
against the CPU oracle (oracle/nmr.py + oracle/csrc/nmr_raster.c) on identical inputs.

Bar: face_index and alpha bit-exact (integer / multiples of 1/4); gradients within 1e-4 of the
oracle's gradient scale (fp32 sums in a different order)."""
import numpy as np
import pytest
import torch

from homan_b200 import synth

pytestmark = pytest.mark.gpu


def _scene(T, obj, seed, mesh="obj"):
    clip = synth.make_clip(T, obj, seed=seed)
    if mesh == "obj":
        return clip["gt"]["verts_obj"], clip["obj_faces"], clip["K_roi_obj"]
    return clip["gt"]["verts_hand"], clip["asset"]["f"], clip["K_roi_hand"]


def _oracle_ndc(verts, K):
    from oracle import nmr
    return nmr.projection(torch.from_numpy(verts), torch.from_numpy(K), torch.eye(3)[None], torch.zeros(1, 3),
                          torch.zeros(1, 5), 1)


def _oracle_render(ndc, faces, image_size, aa, grad_alpha=None):
    from oracle import nmr
    ndc = ndc.clone().requires_grad_(grad_alpha is not None)
    f = torch.from_numpy(faces.astype(np.int64))[None].repeat(ndc.shape[0], 1, 1)
    f2 = torch.cat((f, f[:, :, [2, 1, 0]]), dim=1)
    fv = nmr.vertices_to_faces(ndc, f2)
    alpha, fi = nmr.rasterize_silhouettes(fv, image_size, aa, return_face_index=True)
    g = None
    if grad_alpha is not None:
        alpha.backward(grad_alpha)
        g = ndc.grad
    return alpha.detach(), fi, g


@pytest.mark.parametrize("noise", [True, False])
@pytest.mark.parametrize("mesh,obj,T,aa", [("obj", "cube", 2, True), ("obj", "ellipsoid500", 3, True),
                                           ("hand", "ellipsoid80", 3, True), ("obj", "ellipsoid80", 2, False)])
def test_forward_bit_exact_and_backward(mesh, obj, T, aa, noise):
    """noise=True: per-pixel random gradient (every sweep line overflows the run lists -> bit-line path);
    noise=False: the piecewise-constant gradient of the silhouette loss (run-length / closed-form path)."""
    from homan_b200 import ops
    verts, faces, K = _scene(T, obj, seed=11, mesh=mesh)
    ndc = _oracle_ndc(verts, K).contiguous()
    R = 256
    rng = np.random.default_rng(5)
    # gradient of a silhouette loss against a shifted target (both signs, zeros away from the boundary)
    alpha_ref, fi_ref, _ = _oracle_render(ndc, faces, R, aa)
    target = torch.roll(alpha_ref, shifts=(9, -13), dims=(1, 2))
    target = (target > 0.5).float()
    keep = torch.ones_like(target)
    keep[:, :, 40:70] = 0
    grad_alpha = (2 * keep * (keep * alpha_ref - target) / keep.sum()).float()
    if noise:
        grad_alpha = grad_alpha * torch.from_numpy(rng.uniform(0.5, 1.5, size=grad_alpha.shape).astype(np.float32))
    _, _, g_ref = _oracle_render(ndc, faces, R, aa, grad_alpha)

    ndc_d = ndc.cuda().requires_grad_()
    faces_d = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
    alpha, fi = ops.rasterize_silhouettes(ndc_d, faces_d, R, aa, return_face_index=True)
    n_bad_fi = int((fi.cpu() != fi_ref).sum())
    assert n_bad_fi == 0, f"face_index differs at {n_bad_fi} pixels"
    assert torch.equal(alpha.cpu(), alpha_ref)
    alpha.backward(grad_alpha.cuda())
    g = ndc_d.grad.cpu()
    scale = g_ref.abs().max().item()
    assert scale > 0
    err = (g - g_ref).abs().max().item()
    assert err <= 1e-4 * scale, (err, scale)
    assert g[:, :, 2].abs().max().item() == 0.0


def test_projection_matches_oracle():
    from homan_b200 import ops
    verts, faces, K = _scene(4, "ellipsoid500", seed=3)
    ref = _oracle_ndc(verts, K)
    v = torch.from_numpy(verts).cuda().requires_grad_()
    out = ops.project(v, torch.from_numpy(K).cuda(), orig_size=1.0)
    assert (out.detach().cpu() - ref).abs().max().item() < 2e-6
    w = torch.from_numpy(np.random.default_rng(0).normal(size=ref.shape).astype(np.float32))
    (out * w.cuda()).sum().backward()
    v_ref = torch.from_numpy(verts).requires_grad_()
    from oracle import nmr
    (nmr.projection(v_ref, torch.from_numpy(K), torch.eye(3)[None], torch.zeros(1, 3), torch.zeros(1, 5), 1) * w).sum().backward()
    scale = v_ref.grad.abs().max().item()
    assert (v.grad.cpu() - v_ref.grad).abs().max().item() <= 1e-4 * scale


def test_sil_loss_kernel():
    from homan_b200._lib import call, current_stream, ptr
    rng = np.random.default_rng(1)
    B, R = 3, 256
    alpha = torch.from_numpy(rng.integers(0, 5, size=(B, R, R)).astype(np.float32) / 4)
    target = torch.from_numpy(rng.integers(-1, 2, size=(B, R, R)).astype(np.int8))
    keep, ref = (target >= 0).float(), (target > 0).float()
    norm = torch.tensor([1.0 / keep.sum() / B] * B)
    img = keep * alpha
    loss_ref = ((img - ref) ** 2).sum((1, 2)) * norm
    iou_ref = (img * ref).sum((1, 2)) / ((img + ref).clamp(0, 1).sum((1, 2)) + 1e-6)
    g_ref = 0.7 * 2 * keep * (img - ref) * norm[:, None, None]
    a, t, n = alpha.cuda(), target.cuda(), norm.cuda()
    out = torch.zeros(B, 2, device="cuda")
    g = torch.empty(B, R, R, device="cuda")
    call("hm_sil_loss_fwd_bwd", ptr(a), ptr(t), ptr(n), 0.7, B, R, out.data_ptr(), 2, out.data_ptr() + 4, 2, ptr(g),
         current_stream())
    torch.cuda.synchronize()
    assert torch.allclose(out[:, 0].cpu(), loss_ref, rtol=1e-5)
    assert torch.allclose(out[:, 1].cpu(), iou_ref, rtol=1e-5)
    assert torch.allclose(g.cpu(), g_ref, rtol=1e-6, atol=1e-12)
