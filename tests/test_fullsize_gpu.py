"""GPU: BASELINE.json's full-size configurations through size-independent properties (the CPU oracle would
take minutes to hours there), plus a one-image oracle comparison of the 20 k-face stress mesh."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _render_setup(cfg, n_images=None):
    from homan_b200 import ops, synth
    from homan_b200.workload import CONFIGS
    c = CONFIGS[cfg]
    clip = synth.make_clip(c["T"], c["obj"], seed=c["seed"])
    inits = synth.make_inits(clip, c["P"], seed=c["seed"])
    P, T = c["P"], c["T"]
    R = np.einsum("vk,ptkj->ptvj", clip["obj_verts_can"], inits["obj_R"].astype(np.float32))
    verts = (R + inits["obj_t"][:, :, None]).reshape(P * T, -1, 3).astype(np.float32)
    K = np.tile(clip["K_roi_obj"][None], (P, 1, 1, 1)).reshape(P * T, 3, 3)
    if n_images:
        verts, K = verts[:n_images], K[:n_images]
    return ops, torch.from_numpy(verts).cuda(), torch.from_numpy(K).cuda(), clip["obj_faces"]


@pytest.mark.parametrize("cfg", ["cfg3", "cfg5"])
def test_raster_properties_at_full_size(cfg):
    ops, verts, K, faces = _render_setup(cfg)
    B, F = verts.shape[0], faces.shape[0]
    f = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
    ndc = ops.project(verts, K, orig_size=1.0).requires_grad_()
    alpha, fi = ops.rasterize_silhouettes(ndc, f, 256, True, return_face_index=True)
    assert alpha.shape == (B, 256, 256) and fi.shape == (B, 512, 512)
    q = alpha * 4
    assert torch.equal(q, q.round()) and float(alpha.min()) >= 0 and float(alpha.max()) <= 1
    assert int(fi.min()) >= -1 and int(fi.max()) < 2 * F
    # alpha is the flipped 2x2 pooled coverage of face_index
    cov = (fi >= 0).float().flip(1)
    assert torch.equal(torch.nn.functional.avg_pool2d(cov[:, None], 2)[:, 0], alpha)
    # a closed mesh in front of the camera covers a plausible part of its ROI crop
    frac = cov.mean().item()
    assert 0.15 < frac < 0.8, frac
    # determinism of the forward (z ties resolve to the lowest face index, no race)
    alpha2, fi2 = ops.rasterize_silhouettes(ndc.detach(), f, 256, True, return_face_index=True)
    assert torch.equal(fi, fi2)
    # backward: linear in grad_alpha, nothing on z, finite
    target = torch.roll(alpha.detach(), shifts=(7, -11), dims=(1, 2)).round()
    g = 2 * (alpha.detach() - target) / target[0].numel()
    (g1,) = torch.autograd.grad(alpha, ndc, g, retain_graph=True)
    (g2,) = torch.autograd.grad(alpha, ndc, 2 * g)
    assert torch.isfinite(g1).all() and float(g1[:, :, 2].abs().max()) == 0.0
    scale = float(g1.abs().max())
    assert scale > 0 and float((g2 - 2 * g1).abs().max()) <= 1e-4 * scale


def test_stress_mesh_matches_oracle_on_one_image():
    from oracle import nmr
    ops, verts, K, faces = _render_setup("cfg5", n_images=1)
    ndc = ops.project(verts, K, orig_size=1.0)
    f = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
    ndc_d = ndc.detach().clone().requires_grad_()
    alpha, fi = ops.rasterize_silhouettes(ndc_d, f, 256, True, return_face_index=True)
    fl = torch.from_numpy(faces.astype(np.int64))[None]
    ndc_c = ndc.detach().cpu().requires_grad_()
    fv = nmr.vertices_to_faces(ndc_c, torch.cat((fl, fl[:, :, [2, 1, 0]]), 1))
    a_ref, fi_ref = nmr.rasterize_silhouettes(fv, 256, True, return_face_index=True)
    assert int((fi.cpu() != fi_ref).sum()) == 0 and torch.equal(alpha.cpu(), a_ref.detach())
    target = torch.roll(a_ref.detach(), shifts=(5, 9), dims=(1, 2)).round()
    g = 2 * (a_ref.detach() - target) / target.numel()
    a_ref.backward(g)
    alpha.backward(g.cuda())
    scale = ndc_c.grad.abs().max().item()
    assert (ndc_d.grad.cpu() - ndc_c.grad).abs().max().item() <= 1e-4 * scale


def test_engine_at_cfg3_size_is_finite_improves_and_is_problem_separable(mano_assets):
    from homan_b200.engine import FitEngine
    from homan_b200.workload import make_workload
    batch, lw = make_workload("cfg3", mano_asset=mano_assets["right"])
    eng = FitEngine(batch, lw, mano_asset=mano_assets["right"], use_graph=True)
    out = eng.fit(30)
    tot = out["total"]
    assert tot.shape == (30, 16) and np.isfinite(tot).all()
    # Adam with lr 0.1 on rot6d / PCA overshoots on some inits early on (the reference's own curves do:
    # tests/golden ref_small_step2 goes 0.39 -> 0.76 -> 0.77 -> 0.61); the best init must improve
    assert tot[-1].min() < tot[0].min(), (tot[0], tot[-1])
    for k, v in out["params"].items():
        assert np.isfinite(v).all(), k
    bi, bl = eng.best_init(clips=1)
    assert int(bi[0]) == int(np.argmin(tot[-1]))
    # problem 5 alone follows the same first iterations (per-problem normalisers, no cross-talk)
    sub = {k: (v[5:6] if isinstance(v, np.ndarray) and v.shape[:1] == (16,) and k not in
               ("obj_verts_can", "obj_faces", "hand_faces") else v) for k, v in batch.items()}
    one = FitEngine(sub, lw, mano_asset=mano_assets["right"], use_graph=False).fit(2)
    assert np.allclose(one["total"][:, 0], tot[:2, 5], rtol=1e-4)
