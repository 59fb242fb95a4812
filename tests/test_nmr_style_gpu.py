"""GPU: the NMR-style comparator (baseline/nmr_style, scalar kernels organised like the upstream extension)
reproduces the oracle's semantics, so timing it is timing "the reference's neural_renderer path" as specified."""
import numpy as np
import pytest
import torch

from homan_b200 import synth

pytestmark = pytest.mark.gpu


def test_comparator_matches_oracle():
    from baseline import nmr_style
    from oracle import nmr
    clip = synth.make_clip(2, "ellipsoid80", seed=9)
    v, f, K = clip["gt"]["verts_obj"], clip["obj_faces"], clip["K_roi_obj"]
    ndc = nmr.projection(torch.from_numpy(v), torch.from_numpy(K), torch.eye(3)[None], torch.zeros(1, 3),
                         torch.zeros(1, 5), 1).contiguous()
    fl = torch.from_numpy(f.astype(np.int64))[None].repeat(2, 1, 1)
    ndc_c = ndc.clone().requires_grad_()
    fv = nmr.vertices_to_faces(ndc_c, torch.cat((fl, fl[:, :, [2, 1, 0]]), 1))
    a_ref, fi_ref = nmr.rasterize_silhouettes(fv, 256, True, return_face_index=True)
    target = torch.roll(a_ref.detach(), shifts=(6, -8), dims=(1, 2)).round()
    g = 2 * (a_ref.detach() - target) / target.numel()
    a_ref.backward(g)
    ndc_d = ndc.cuda().requires_grad_()
    a, fi = nmr_style.render_silhouettes(ndc_d, fl.cuda(), 256, True, return_face_index=True)
    assert int((fi.cpu() != fi_ref).sum()) == 0 and torch.equal(a.cpu(), a_ref.detach())
    a.backward(g.cuda())
    scale = ndc_c.grad.abs().max().item()
    assert (ndc_d.grad.cpu() - ndc_c.grad).abs().max().item() <= 1e-4 * scale
    # the face-parallel forward (organisation of the fork the reference installs) gives the same maps
    a2, fi2 = nmr_style.render_silhouettes(ndc.cuda(), fl.cuda(), 256, True, return_face_index=True, fast=True)
    assert torch.equal(fi2, fi) and torch.equal(a2, a.detach())
