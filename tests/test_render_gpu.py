"""GPU: forward of the RGB / depth render used by the reference for visualisation (`nr.Renderer.render`,
/root/reference/homan/homan.py:510-613) through the neural_renderer drop-in, against the CPU
oracle (oracle/nmr.py::Renderer.render, restated from the un-vendored package: parity unpinned). Bar: alpha and depth
bit-exact (same z-buffer arithmetic), colours 1e-5 (flat lighting: hm_face_lighting on the GPU, PyTorch in the oracle)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(T=3, seed=0):
    from homan_b200 import synth
    clip = synth.make_clip(T, "ellipsoid500", seed=seed)
    vo, fo = clip["gt"]["verts_obj"], clip["obj_faces"]
    vh, fh = clip["gt"]["verts_hand"], clip["asset"]["f"]
    verts = np.concatenate((vo, vh), axis=1).astype(np.float32)
    faces = np.concatenate((fo, fh + vo.shape[1])).astype(np.int32)
    rng = np.random.default_rng(seed)
    colours = rng.uniform(0.1, 1.0, size=(faces.shape[0], 3)).astype(np.float32)
    return verts, faces, colours, clip["K_roi_obj"].astype(np.float32)


@pytest.mark.parametrize("aa", [True, False])
def test_render_matches_oracle(aa):
    from homan_b200.shims import neural_renderer as nr_gpu
    from oracle import nmr
    verts, faces, colours, K = _scene()
    T, F = verts.shape[0], faces.shape[0]
    tex = torch.from_numpy(colours).view(1, F, 1, 1, 1, 3).repeat(T, 1, 1, 1, 1, 1)
    kw = dict(image_size=128, anti_aliasing=aa, orig_size=1, background_color=(0.2, 0.4, 0.6),
              light_intensity_ambient=0.4, light_intensity_directional=0.6, light_direction=(0.3, 0.8, -0.5))
    ref = nmr.Renderer(K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3), **kw)
    rgb_r, depth_r, alpha_r = ref.render(torch.from_numpy(verts), torch.from_numpy(faces)[None].repeat(T, 1, 1), tex)
    dev = "cuda"
    got = nr_gpu.Renderer(K=torch.from_numpy(K).to(dev), R=torch.eye(3, device=dev)[None], t=torch.zeros(1, 3, device=dev), **kw)
    rgb, depth, alpha = got.render(torch.from_numpy(verts).to(dev), torch.from_numpy(faces).to(dev)[None].repeat(T, 1, 1),
                                   tex.to(dev))
    assert rgb.shape == (T, 3, 128, 128) and depth.shape == (T, 128, 128)
    assert torch.equal(alpha.cpu(), alpha_r)
    assert torch.equal(depth.cpu(), depth_r)
    assert float((rgb.cpu() - rgb_r).abs().max()) <= 1e-5
    assert 0.05 < float(alpha_r.mean()) < 0.9 and float(depth_r.min()) < 1.0 and float(depth_r.max()) == 100.0
    # mode=None of __call__ is the same render
    rgb2, depth2, alpha2 = got(torch.from_numpy(verts).to(dev), torch.from_numpy(faces).to(dev)[None].repeat(T, 1, 1),
                               tex.to(dev))
    assert torch.equal(rgb2, rgb) and torch.equal(depth2, depth) and torch.equal(alpha2, alpha)
