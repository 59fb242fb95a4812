"""Helpers to load the golden fixtures produced by scripts/make_golden.py."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAMS = ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
          "mano_pca_pose", "mano_rot", "mano_trans", "mano_betas")


def load(name, mano_asset):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    batch["target_masks_object"] = batch["target_masks_object"].astype(np.float32)
    batch["target_masks_hand"] = batch["target_masks_hand"].astype(np.float32)
    batch.update(P=batch["obj_t"].shape[0], T=batch["obj_t"].shape[1], side="right", image_size=640,
                 mano_asset=mano_asset)
    lw = {str(k): float(v) for k, v in zip(z["lw_keys"], z["lw_vals"])}
    return z, batch, lw, int(z["iters"])
