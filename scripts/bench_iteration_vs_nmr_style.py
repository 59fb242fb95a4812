"""COMPARATOR (measurement infrastructure, not product): one optimizer iteration of the reference loop organised as the
reference runs it on a GPU, against this repo's fused iteration, on the same B200 and the same cfg3 problem.

The reference's own GPU path cannot be built here (its renderer, SDF and MANO packages are un-vendored, no network), so
it is re-created from the pieces the repo already has:

  * the loop, HOMan.forward, the losses and Adam: `oracle/homan_ref.py::ClipModel` (the torch restatement of
    /root/reference/homan/jointopt.py:128-192 and homan/homan.py:421-508 that the goldens pin) moved to the GPU - eager
    PyTorch, one clip of T frames and ONE init per run, as `optimize_hand_object` runs it;
  * `neural_renderer`: `baseline/nmr_style` (thread-per-face `backward_pixel_map` with full sweeps, forward face-parallel
    over the pixel bounding box as in the fork the reference installs, doubling / gather / flip / avg-pool in PyTorch);
  * `sdf.SDF`: dense 32^3 grids, one thread per voxel over all faces (`hm_sdf_grid`, organised like that extension),
    four per iteration: two for the collision term and two that `compute_contact_loss` builds and never uses
    (/root/reference/homan/interactions/contactloss.py:169, scenesdf.py:112-147), then `F.grid_sample`.

The reference fits the P random inits one after the other, so a whole-batch iteration of the comparator costs P single-init
iterations; ours is one graph replay over all P x T images.  Prints one JSON object.

    python scripts/bench_iteration_vs_nmr_style.py [--workload cfg3] [--iters 10] [--out file]
"""
import argparse
import inspect
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import nmr_style  # noqa: E402
from homan_b200 import workload  # noqa: E402
from homan_b200.engine import FitEngine  # noqa: E402
from homan_b200.shims.sdf import SDF as DenseSDF  # noqa: E402
from oracle import homan_ref, nmr as oracle_nmr  # noqa: E402


def _gpu_rasterize_silhouettes(faces, image_size=256, anti_aliasing=True, near=oracle_nmr.DEFAULT_NEAR,
                               far=oracle_nmr.DEFAULT_FAR, eps=oracle_nmr.DEFAULT_EPS, return_face_index=False):
    """oracle.nmr.rasterize_silhouettes on the NMR-style kernels: faces [B,nf,3,3] (already doubled)."""
    S = image_size * 2 if anti_aliasing else image_size
    B, nf = faces.shape[:2]
    alpha, fi = nmr_style._Rasterize.apply(faces.reshape(B, nf, 9), S, near, far, eps, True)
    alpha = alpha.flip(1)
    if anti_aliasing:
        alpha = torch.nn.functional.avg_pool2d(alpha[:, None], 2)[:, 0]
    return (alpha, fi) if return_face_index else alpha


def _to_cuda(obj, seen=None):
    """Moves every tensor attribute of the oracle's model objects to the GPU (leaf parameters stay leaves)."""
    seen = set() if seen is None else seen
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.nn.Module):
        obj.cuda()   # (registered buffers; plain tensor attributes below)
    for k, v in list(vars(obj).items()):
        if torch.is_tensor(v):
            rg = v.requires_grad
            setattr(obj, k, v.detach().cuda().requires_grad_(rg))
        elif isinstance(v, dict):
            for kk, vv in list(v.items()):
                if torch.is_tensor(vv):
                    v[kk] = vv.cuda()
                elif isinstance(vv, torch.nn.Module):
                    _to_cuda(vv, seen)
        elif hasattr(v, "__dict__") and not isinstance(v, type) and not inspect.isroutine(v):
            _to_cuda(v, seen)


class _Model(homan_ref.ClipModel):
    """ClipModel + the two dense grids the reference's contact term builds and discards."""

    def contact(self, hand, obj):
        with torch.no_grad():
            homan_ref.sdf_scene([hand.detach(), obj.detach()], [self.closed_faces, self.faces_object[0]])
        return super().contact(hand, obj)


def stacked_problem(batch):
    """All P inits of the clip as ONE clip of P x T frames (not how the reference runs - it fits one init per call -
    but the most favourable batching of its organisation: every kernel sees all 480 images at once)."""
    prob = homan_ref.problem_slice(batch, 0)
    for k in ("obj_t", "obj_R", "hand_t", "hand_R", "pca", "mano_rot", "mano_trans", "betas", "target_masks_object",
              "target_masks_hand", "K_roi_obj", "K_roi_hand", "verts2d"):
        if k in batch:
            a = np.asarray(batch[k])
            prob[k] = a.reshape((a.shape[0] * a.shape[1],) + a.shape[2:])
    cam = np.asarray(batch["camintr"])[0]
    if cam.ndim == 3 and cam.shape[0] > 1:   # one intrinsic matrix per frame
        prob["camintr"] = np.asarray(batch["camintr"]).reshape((-1,) + cam.shape[1:])
    return prob


def time_comparator(batch, lw, iters, warmup=2, stacked=False):
    # the oracle modules run on whatever device their tensors live on once the two native calls are replaced
    homan_ref.nmr.rasterize_silhouettes = _gpu_rasterize_silhouettes
    oracle_nmr.rasterize_silhouettes = _gpu_rasterize_silhouettes
    homan_ref.sdfmod.SDF = DenseSDF
    torch.set_default_device("cuda")   # the restatement creates a few constants inline (torch.eye, torch.zeros)
    mano = homan_ref.ManoPca(batch["mano_asset"])
    model = _Model(stacked_problem(batch) if stacked else homan_ref.problem_slice(batch, 0), mano,
                   batch["mano_asset"]["closed_faces"])
    _to_cuda(mano)
    _to_cuda(model)
    opt = homan_ref.make_optimizer(model, 1e-2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    parts = {}

    def one():
        opt.zero_grad()
        losses, _ = model.forward(lw)
        total = sum(v * lw[k.replace("loss", "lw")] for k, v in losses.items())
        total.backward()
        opt.step()
        return float(total)

    for _ in range(warmup):
        first = one()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        last = one()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    # share of the native pieces (timed alone, same inputs)
    with torch.no_grad():
        vo, _ = model.verts_object()
        vh, _ = model.verts_hand()
    for name, fn in (("sdf_4_dense_grids", lambda: [homan_ref.sdf_scene([vh, vo], [model.closed_faces, model.faces_object[0]])
                                                     for _ in range(2)]),):
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(3):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        parts[name] = ev[0].elapsed_time(ev[1]) / 3
    return ms, parts, first, last


def time_ours(batch, lw, asset, iters, warmup=5):
    eng = FitEngine(batch, lw, lr=1e-2, mano_asset=asset, use_graph=True)
    eng.capture()
    for _ in range(warmup):
        eng.step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        eng.step()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from homan_b200 import synth
    asset = synth.make_mano_asset(0, "right")
    batch, lw = workload.make_workload(a.workload, mano_asset=asset)
    batch = dict(batch)
    batch.setdefault("mano_asset", asset)
    P, T = int(np.asarray(batch["obj_t"]).shape[0]), int(np.asarray(batch["obj_t"]).shape[1])
    ours_ms = time_ours(batch, lw, asset, max(a.iters * 5, 50))
    ref_ms, parts, l0, l1 = time_comparator(batch, lw, a.iters)
    st_ms, st_parts, _, _ = time_comparator(batch, lw, max(a.iters // 2, 2), stacked=True)
    out = {"workload": a.workload, "inits": P, "frames": T,
           "comparator_ms_per_iteration_one_init": ref_ms,
           "comparator_ms_per_whole_batch_iteration": ref_ms * P,
           "comparator_parts_ms_one_init": parts,
           "comparator_loss_first_last": [l0, l1],
           "comparator_stacked_ms_per_whole_batch_iteration": st_ms,
           "comparator_stacked_parts_ms": st_parts,
           "ours_ms_per_whole_batch_iteration": ours_ms,
           "speedup_iteration": ref_ms * P / ours_ms,
           "speedup_iteration_vs_stacked": st_ms / ours_ms,
           "what": "eager-PyTorch reference loop (oracle.homan_ref.ClipModel on the GPU) + NMR-style renderer kernels + "
                   "4 dense 32^3 SDF grids per iteration, one init at a time (x inits) vs one fused graph replay over "
                   "all inits x frames; same B200, same problem. 'stacked': the same eager loop given all inits as one clip of "
                   "inits x frames images (the most favourable batching of the reference's organisation)"}
    s = json.dumps(out)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
