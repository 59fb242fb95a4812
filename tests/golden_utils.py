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


def reference_inputs(batch, p, mano_asset):
    """Problem p of a golden batch in the argument schema of the reference's optimize_hand_object
    (lists of per-frame dicts)."""
    import torch
    T = batch["T"]
    t32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))  # noqa: E731
    person, obj = [], []
    faces_hand = torch.from_numpy(np.asarray(batch["hand_faces"]).astype(np.int32))[None]
    for t in range(T):
        person.append({
            "translations": t32(batch["hand_t"][p, t]).view(1, 1, 3), "rotations": t32(batch["hand_R"][p, t]).view(1, 3, 3),
            "hand_side": ["right"], "faces": faces_hand,
            "mano_trans": t32(batch["mano_trans"][p, t]).view(1, 3), "mano_rot": t32(batch["mano_rot"][p, t]).view(1, 3),
            "mano_betas": t32(batch["betas"][p, t]).view(1, 10), "mano_pca_pose": t32(batch["pca"][p, t]).view(1, -1),
            "target_masks": t32(batch["target_masks_hand"][p, t]).view(1, 256, 256),
            "masks": torch.zeros(1, 8, 8, dtype=torch.bool), "verts": torch.zeros(1, 778, 3),
            "verts2d": t32(batch["verts2d"][p, t]).view(1, 778, 2), "K_roi": t32(batch["K_roi_hand"][p, t]).view(1, 3, 3),
            "cams": torch.ones(1, 3)})
        obj.append({
            "translations": t32(batch["obj_t"][p, t]).view(1, 1, 3), "rotations": t32(batch["obj_R"][p, t]).view(1, 3, 3),
            "target_masks": t32(batch["target_masks_object"][p, t]).view(1, 256, 256),
            "full_mask": torch.zeros(8, 8, dtype=torch.bool), "K_roi": t32(batch["K_roi_obj"][p, t]).view(1, 1, 3, 3)})
    return dict(person_parameters=person, object_parameters=obj,
                objvertices=np.repeat(np.asarray(batch["obj_verts_can"])[None], T, 0),
                objfaces=np.repeat(np.asarray(batch["obj_faces"])[None], T, 0), camintr=batch["camintr"][p])
