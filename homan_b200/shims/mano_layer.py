"""Drop-in for `mano.model.load(...)` (/root/reference/homan/manomodel.py:9,19-82,110-123): a callable layer
(betas, global_orient, hand_pose, transl) -> (vertices, joints, betas, transl, global_orient, full_pose) with
`.hand_mean` / `.hand_components`, evaluated by the fused LBS kernel (csrc/mano.cu) with its analytic backward."""
import os
import pickle

import numpy as np
import torch
from torch import nn

from .. import engine as _engine
from .._lib import call, current_stream, ptr


def load_asset(path_or_dict, is_right=True):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    path = path_or_dict
    if os.path.isdir(path):
        path = os.path.join(path, "MANO_RIGHT.pkl" if is_right else "MANO_LEFT.pkl")
    with open(path, "rb") as fh:
        data = pickle.load(fh, encoding="latin1")
    out = {k: np.asarray(v.todense() if hasattr(v, "todense") else v) for k, v in dict(data).items()
           if k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "hands_components",
                    "hands_mean", "kintree_table", "parents", "f", "closed_faces")}
    if "parents" not in out:
        parents = out["kintree_table"][0].astype(np.int64).copy()
        parents[0] = -1
        out["parents"] = parents
    return out


class _ManoLBS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, blob, pose45, rot, betas, transl):
        B = pose45.shape[0]
        p, r, b, t = (x.detach().contiguous().float() for x in (pose45, rot, betas, transl))
        verts = torch.empty(B, 778, 3, device=p.device)
        joints = torch.empty(B, 16, 3, device=p.device)
        # the layer takes the 45-D axis-angle pose directly: identity "components", zero mean in the blob
        call("hm_mano_fwd", ptr(blob), 45, 0, ptr(p), 45, ptr(r), ptr(b), ptr(t), None, None, None, B, ptr(verts),
             ptr(joints), None, current_stream())
        ctx.save_for_backward(blob, p, r, b, t)
        ctx.mark_non_differentiable(joints)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, _g_joints):
        blob, p, r, b, t = ctx.saved_tensors
        B = p.shape[0]
        gv = g_verts.detach().contiguous().float()
        gp, gr, gb, gt = (torch.zeros_like(x) for x in (p, r, b, t))
        call("hm_mano_bwd", ptr(blob), 45, 0, ptr(p), 45, ptr(r), ptr(b), ptr(t), None, None, None, B, None, ptr(gv), None,
             ptr(gp), ptr(gr), ptr(gb), ptr(gt), None, None, current_stream())
        return None, gp, gr, gb, gt


class ManoLayer(nn.Module):
    def __init__(self, asset, num_pca_comps=6, use_pca=True, flat_hand_mean=False, is_right=True, batch_size=1):
        super().__init__()
        a = load_asset(asset, is_right)
        self.use_pca, self.num_pca_comps, self.flat_hand_mean, self.is_right = use_pca, num_pca_comps, flat_hand_mean, is_right
        comps = torch.as_tensor(np.asarray(a["hands_components"]), dtype=torch.float32)[:num_pca_comps]
        mean = torch.zeros(45) if flat_hand_mean else torch.as_tensor(np.asarray(a["hands_mean"]), dtype=torch.float32)
        self.register_buffer("hand_components", comps.cuda())
        self.register_buffer("hand_mean", mean.cuda())
        self.register_buffer("pose_mean", torch.cat([torch.zeros(3), mean]).cuda())
        self.register_buffer("faces_tensor", torch.as_tensor(np.asarray(a["f"]).astype(np.int64)).cuda())
        ident = dict(a)
        ident["hands_components"] = np.eye(45, dtype=np.float32)
        ident["hands_mean"] = np.zeros(45, dtype=np.float32)
        self.register_buffer("blob", _engine.mano_blob(ident, 45, "cuda"))

    def forward(self, betas=None, global_orient=None, hand_pose=None, transl=None, **kwargs):
        if self.use_pca:
            hand_pose = hand_pose @ self.hand_components
        hand_pose = hand_pose + self.hand_mean
        if transl is None:
            transl = torch.zeros_like(global_orient)
        verts, joints = _ManoLBS.apply(self.blob, hand_pose, global_orient, betas, transl)
        full_pose = torch.cat([global_orient, hand_pose], 1)
        return verts, joints, betas, transl, global_orient, full_pose


def load(model_path=None, is_right=True, model_type="mano", num_pca_comps=6, use_pca=True, batch_size=1,
         flat_hand_mean=False, **kwargs):
    return ManoLayer(model_path, num_pca_comps=num_pca_comps, use_pca=use_pca, flat_hand_mean=flat_hand_mean,
                     is_right=is_right, batch_size=batch_size)
