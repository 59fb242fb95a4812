"""Generates tests/golden/*.npz by running the UNMODIFIED reference Python
(/root/reference/homan/jointopt.py::optimize_hand_object) on CPU under oracle/refshim.py.

Only runnable in the build container (needs /root/reference).  The fixtures hold the complete
synthetic inputs (so they do not depend on homan_b200/synth.py staying bit-stable) and the
reference outputs: per-iteration loss / metric values, parameter gradients after the first
backward, and the fitted parameters.

    python scripts/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from homan_b200 import synth  # noqa: E402
from oracle import nmr, refshim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (T, object, P, seed, loss weights, iterations)
    "ref_cfg1_cube": dict(T=1, obj="cube", P=1, seed=1000, lw=synth.default_loss_weights(lw_sil_obj=1.0), iters=12),
    "ref_small_step1": dict(T=5, obj="ellipsoid80", P=2, seed=2000, lw=synth.step1_loss_weights(), iters=4),
    "ref_small_step2": dict(T=4, obj="ellipsoid80", P=2, seed=3000, lw=synth.step2_loss_weights(), iters=4),
}

BATCH_KEYS = ("obj_verts_can", "obj_faces", "hand_faces", "camintr", "K_roi_obj", "K_roi_hand", "verts2d",
              "target_masks_object", "target_masks_hand", "obj_R", "obj_t", "hand_R", "hand_t", "pca",
              "mano_rot", "mano_trans", "betas")


def render_fn(verts, faces, K):
    r = nmr.Renderer(image_size=256, K=torch.from_numpy(K), R=torch.eye(3)[None], t=torch.zeros(1, 3),
                     orig_size=1, anti_aliasing=False)
    f = torch.from_numpy(faces.astype(np.int32))[None].repeat(len(verts), 1, 1)
    return r(torch.from_numpy(verts), f, mode="silhouettes").numpy()


def main():
    assert refshim.reference_available(), "needs /root/reference"
    os.makedirs(GOLDEN, exist_ok=True)
    assets = {"right": synth.make_mano_asset(0, "right"), "left": synth.make_mano_asset(1, "left")}
    scratch = tempfile.mkdtemp(prefix="homan_golden_")
    refshim.install(scratch, assets)
    for name, c in CASES.items():
        clip = synth.make_clip(c["T"], c["obj"], seed=c["seed"], mano_asset=assets["right"], render_fn=render_fn)
        inits = synth.make_inits(clip, c["P"], seed=c["seed"])
        batch = synth.make_batch(clip, inits)
        out = {"in_" + k: batch[k] for k in BATCH_KEYS}
        out["in_target_masks_object"] = batch["target_masks_object"].astype(np.int8)
        out["in_target_masks_hand"] = batch["target_masks_hand"].astype(np.int8)
        out["lw_keys"] = np.array(sorted(c["lw"]))
        out["lw_vals"] = np.array([c["lw"][k] for k in sorted(c["lw"])], dtype=np.float64)
        out["iters"] = np.array(c["iters"])
        out["mano_seed_right"] = np.array(0)
        for p in range(c["P"]):
            inp = synth.reference_inputs(clip, inits, p)
            # gradients of the first backward (lr = 0 keeps the parameters in place)
            model, ev = refshim.run_reference_fit(inp, c["lw"], 1, scratch, lr=0.0)
            for k, v in model.named_parameters():
                if v.grad is not None and "cams" not in k:
                    out[f"grad0_{k}_p{p}"] = v.grad.numpy().copy()
            with refshim.record_trajectory() as rec:
                model, ev = refshim.run_reference_fit(inp, c["lw"], c["iters"], scratch, lr=1e-2)
            assert len(rec.snapshots) == c["iters"]
            for k in ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
                      "mano_pca_pose", "mano_betas"):
                out[f"traj_{k}_p{p}"] = np.stack([snap[k] for snap in rec.snapshots])
            for k, v in ev.items():
                out[f"ev_{k}_p{p}"] = np.asarray(v, dtype=np.float64)
            sd = model.state_dict()
            for k in ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
                      "mano_pca_pose", "mano_rot", "mano_trans", "mano_betas"):
                out[f"final_{k}_p{p}"] = sd[k].numpy().copy()
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", {k: float(ev[k][0]) for k in ev if k.startswith("loss")})


if __name__ == "__main__":
    main()
