"""Synthetic benchmark / test workloads on the GPU (BASELINE.json configs).

Targets (object / hand masks) are rendered with the product rasteriser itself (anti-aliasing off ->
binary masks), so nothing here touches the CPU oracle.
"""
import numpy as np
import torch

from . import ops, synth

CONFIGS = {
    # name: P inits, T frames, object mesh, loss weights
    "cfg1": dict(P=1, T=1, obj="cube", lw="sil_obj", seed=1000),
    "cfg2": dict(P=16, T=10, obj="ellipsoid500", lw="step1+sil_hand", seed=2000),
    "cfg3": dict(P=16, T=30, obj="ellipsoid500", lw="step2+sil_hand", seed=3000),
    # cfg4: 64 clips x 16 inits x 10 frames over 8 GPUs = 8 whole clips (128 problems, 1280 images) per GPU, every clip
    # with its own object; `clips` = clips per GPU
    "cfg4": dict(P=16, T=10, obj="ellipsoid500@", lw="step1+sil_hand", seed=4000, clips=8, total_clips=64),
    "cfg5": dict(P=32, T=30, obj="ellipsoid20k", lw="sil_obj", seed=5000),
    "tiny": dict(P=2, T=4, obj="ellipsoid80", lw="step2+sil_hand", seed=7000),
    "tiny4": dict(P=2, T=3, obj="ellipsoid80@", lw="step2+sil_hand", seed=7400, clips=3, total_clips=3),
}


def loss_weights(kind):
    if kind == "sil_obj":
        return synth.default_loss_weights(lw_sil_obj=1.0)
    base = synth.step2_loss_weights() if kind.startswith("step2") else synth.step1_loss_weights()
    if kind.endswith("+sil_hand"):
        base["lw_sil_hand"] = 1.0
    return base


def gpu_render_fn(device="cuda"):
    def render(verts, faces, K):
        v = torch.from_numpy(np.ascontiguousarray(verts, dtype=np.float32)).to(device)
        k = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float32)).to(device)
        f = torch.from_numpy(np.ascontiguousarray(faces).astype(np.int32)).to(device)[None]
        ndc = ops.project(v, k, orig_size=1.0)
        out = ops.rasterize_silhouettes(ndc, f, 256, anti_aliasing=False)
        return out.cpu().numpy()
    return render


def make_workload(name, clip_index=0, device="cuda", mano_asset=None, init_shard=0, hand_mesh="delaunay", clips=None):
    """-> (batch dict of numpy arrays [P,T,...], loss weights). `clip_index` selects the synthetic clip (ground-truth
    trajectory and target masks), `init_shard` the block of P random initialisations of that clip, `hand_mesh` the
    triangulation of the synthetic hand when no asset is passed (synth._hand_template). Multi-clip configurations
    (cfg4): `clips` = the clip ids to stack (default: the first CONFIGS[name]["clips"]); the batch is clip-major and
    every clip carries its own object mesh."""
    cfg = CONFIGS[name]
    asset = mano_asset if mano_asset is not None else synth.make_mano_asset(0, "right", mesh=hand_mesh)

    def one(clip_id, obj):
        seed = cfg["seed"] + clip_id
        clip = synth.make_clip(cfg["T"], obj, seed=seed, mano_asset=asset, render_fn=gpu_render_fn(device))
        inits = synth.make_inits(clip, cfg["P"], seed=seed + 7919 * init_shard)
        return synth.make_batch(clip, inits)

    if "clips" in cfg:
        ids = list(range(cfg["clips"])) if clips is None else list(clips)
        batch = synth.concat_batches([one(c, cfg["obj"] + str(cfg["seed"] + c)) for c in ids])
        batch["clip_ids"] = np.asarray(ids, np.int64)
        return batch, loss_weights(cfg["lw"])
    return one(clip_index, cfg["obj"]), loss_weights(cfg["lw"])
