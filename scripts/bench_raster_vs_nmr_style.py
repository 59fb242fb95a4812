"""Silhouette render + backward of BASELINE.json cfg3's meshes on one B200: homan_b200 kernels vs the NMR-style
comparator (baseline/nmr_style: scalar kernels organised like the upstream neural_renderer extension + the
eager PyTorch glue upstream uses; forward in both upstream organisations: pixel-parallel over all faces, and
face-parallel over the pixel bounding box as in the fork the reference installs). Same batch (480 images) on both
sides. Prints one JSON object, including an iteration-level lower bound of the speed-up. Both sides get identical NDC vertices and the same
silhouette-loss gradient; times are CUDA events on the launching stream, per call over all images."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from baseline import nmr_style  # noqa: E402
from homan_b200 import ops, synth  # noqa: E402
from homan_b200.workload import CONFIGS  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    c = CONFIGS["cfg3"]
    asset = synth.make_mano_asset(0, "right")
    clip = synth.make_clip(c["T"], c["obj"], seed=c["seed"], mano_asset=asset)
    inits = synth.make_inits(clip, c["P"], seed=c["seed"])
    P, T = c["P"], c["T"]
    out = {"workload": "cfg3 meshes, 256^2 output / 512^2 raster, anti-aliasing", "images_ours": P * T}
    n_cmp = P * T  # the same batch on both sides (a thread-per-face backward needs the whole batch to fill the GPU)
    out["images_nmr_style"] = n_cmp
    for name in ("object", "hand"):
        if name == "object":
            R = np.einsum("vk,ptkj->ptvj", clip["obj_verts_can"], inits["obj_R"].astype(np.float32))
            verts = (R + inits["obj_t"][:, :, None]).reshape(P * T, -1, 3).astype(np.float32)
            faces, K = clip["obj_faces"], np.tile(clip["K_roi_obj"][None], (P, 1, 1, 1)).reshape(P * T, 3, 3)
        else:
            verts = np.tile(clip["gt"]["verts_hand"][None], (P, 1, 1, 1)).reshape(P * T, 778, 3)
            verts = verts + np.repeat(np.random.default_rng(0).normal(size=(P, 1, 1, 3)) * 0.01, T, 1).reshape(P * T, 1, 3).astype(np.float32)
            faces, K = asset["f"], np.tile(clip["K_roi_hand"][None], (P, 1, 1, 1)).reshape(P * T, 3, 3)
        v, Kd = torch.from_numpy(verts.astype(np.float32)).cuda(), torch.from_numpy(K.astype(np.float32)).cuda()
        f32 = torch.from_numpy(faces.astype(np.int32)).cuda()[None]
        ndc = ops.project(v, Kd, orig_size=1.0).detach()
        B, V, F = ndc.shape[0], ndc.shape[1], faces.shape[0]
        buf = ops.RasterBuffers(B, V, F, 256, True, ndc.device)
        ops.raster_forward(buf, ndc, f32)
        target = torch.roll(buf.alpha, shifts=(5, -7), dims=(1, 2)).round()
        g = (2 * (buf.alpha - target) / (256 * 256 * T)).contiguous()
        gn = torch.zeros(B, V, 3, device=ndc.device)
        t_fwd = timed(lambda: ops.raster_forward(buf, ndc, f32), 5)
        t_bwd = timed(lambda: ops.raster_backward(buf, g, gn.zero_()), 5)
        # comparator on a subset (same images)
        nd = ndc[:n_cmp].clone().requires_grad_()
        fl = f32.long().repeat(n_cmp, 1, 1)
        holder = {}

        def cmp_fwd():
            holder["a"] = nmr_style.render_silhouettes(nd, fl, 256, True)

        def cmp_fwd_fast():
            holder["a"] = nmr_style.render_silhouettes(nd, fl, 256, True, fast=True)

        def cmp_bwd():
            (holder["g"],) = torch.autograd.grad(holder["a"], nd, g[:n_cmp], retain_graph=True)

        c_fwd_pixel = timed(cmp_fwd, 1)
        c_fwd = timed(cmp_fwd_fast, 2)
        c_bwd = timed(cmp_bwd, 1)
        # same answers
        assert torch.equal(holder["a"].detach(), buf.alpha[:n_cmp])
        gn.zero_()
        ops.raster_backward(buf, g, gn)
        scale = holder["g"].abs().max().item()
        err = (holder["g"] - gn[:n_cmp]).abs().max().item()
        assert err <= 1e-4 * scale, (name, err, scale)   # the two renderers must agree on all 480 images
        out[name] = {
            "faces": int(F),
            "ours_ms_per_image": {"forward": t_fwd / B, "backward": t_bwd / B},
            "nmr_style_ms_per_image": {"forward": c_fwd / n_cmp, "forward_pixel_parallel": c_fwd_pixel / n_cmp,
                                       "backward": c_bwd / n_cmp},
            "nmr_style_ms_per_call": {"forward": c_fwd, "backward": c_bwd},
            "speedup": {"forward": (c_fwd / n_cmp) / (t_fwd / B), "backward": (c_bwd / n_cmp) / (t_bwd / B),
                        "forward+backward": ((c_fwd + c_bwd) / n_cmp) / ((t_fwd + t_bwd) / B)},
            "grad_max_rel_diff": err / scale,
        }
    # ---- iteration level: the reference's GPU iteration = its renderer (both meshes, forward + backward) + four
    #      brute-force 32^3 SDF grids + eager PyTorch LBS / losses / Adam; the comparator covers the renderer only,
    #      so (renderer-only comparator time) / (our WHOLE iteration) is a lower bound of the iteration speed-up
    from homan_b200.engine import FitEngine
    from homan_b200.workload import make_workload
    batch, lw = make_workload("cfg3", mano_asset=asset)
    eng = FitEngine(batch, lw, mano_asset=asset, use_graph=True)
    for _ in range(5):
        eng.step()
    ours_iter = timed(eng.step, 50)
    cmp_iter = sum(out[k]["nmr_style_ms_per_call"]["forward"] + out[k]["nmr_style_ms_per_call"]["backward"]
                   for k in ("object", "hand"))
    out["iteration"] = {"ours_whole_iteration_ms": ours_iter, "nmr_style_renderer_only_ms": cmp_iter,
                        "speedup_lower_bound": cmp_iter / ours_iter,
                        "note": "comparator = silhouette render + backward of both meshes only; the reference's "
                                "iteration adds 4 dense SDF grids, LBS, losses and Adam in eager PyTorch"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
