"""ORACLE (test infrastructure; never imported by the product).

CPU stand-in for `mano.model.load` (github.com/hassony2/MANO, smplx-style layer) that
the reference uses at /root/reference/homan/manomodel.py:9,19-82,110-123.  The package
and the licensed MANO pickles are absent (README.md:72-90); semantics follow
SURVEY.md Appendix A.5 (parity unpinned).  Assets are dict pickles with the standard
MANO keys (see homan_b200/synth.py::make_mano_asset).
"""
import pickle

import numpy as np
import torch
from torch import nn


def batch_rodrigues(rot_vecs, epsilon=1e-8):
    """[N,3] axis-angle -> [N,3,3]; angle = ||r + 1e-8|| (smplx convention)."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + epsilon, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view((n, 3, 3))
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(dim=0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def _transform_mat(R, t):
    return torch.cat([torch.nn.functional.pad(R, [0, 0, 0, 1]),
                      torch.nn.functional.pad(t, [0, 0, 0, 1], value=1)], dim=2)


def batch_rigid_transform(rot_mats, joints, parents):
    joints = torch.unsqueeze(joints, dim=-1)
    rel_joints = joints.clone()
    rel_joints[:, 1:] = rel_joints[:, 1:] - joints[:, parents[1:]]
    transforms_mat = _transform_mat(rot_mats.reshape(-1, 3, 3), rel_joints.reshape(-1, 3, 1)).reshape(
        -1, joints.shape[1], 4, 4)
    chain = [transforms_mat[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[parents[i]], transforms_mat[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_homogen = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - torch.nn.functional.pad(torch.matmul(transforms, joints_homogen),
                                                          [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """smplx-style linear blend skinning. pose [B, 48] axis-angle. Returns verts [B,778,3], joints [B,16,3]."""
    B = betas.shape[0]
    v_shaped = v_template + torch.einsum("bl,mkl->bmk", betas, shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)
    ident = torch.eye(3, dtype=betas.dtype, device=betas.device)
    rot_mats = batch_rodrigues(pose.view(-1, 3)).view(B, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:, :, :] - ident).view(B, -1)
    pose_offsets = torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents)
    W = lbs_weights.unsqueeze(dim=0).expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, 16, 16)).view(B, -1, 4, 4)
    homogen_coord = torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype, device=betas.device)
    v_posed_homo = torch.cat([v_posed, homogen_coord], dim=2)
    v_homo = torch.matmul(T, torch.unsqueeze(v_posed_homo, dim=-1))
    return v_homo[:, :, :3, 0], J_transformed


def load_asset(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    with open(path_or_dict, "rb") as fh:
        data = pickle.load(fh, encoding="latin1")
    return {k: np.asarray(v) for k, v in data.items()}


class ManoLayer(nn.Module):
    def __init__(self, asset, num_pca_comps=6, use_pca=True, flat_hand_mean=False, is_right=True,
                 batch_size=1, dtype=torch.float32):
        super().__init__()
        a = load_asset(asset)
        self.num_pca_comps = num_pca_comps
        self.use_pca = use_pca
        self.is_right = is_right
        self.flat_hand_mean = flat_hand_mean
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=dtype)  # noqa: E731
        self.register_buffer("v_template", t(a["v_template"]))
        self.register_buffer("shapedirs", t(a["shapedirs"]))
        self.register_buffer("posedirs", t(a["posedirs"]).reshape(135, -1))
        self.register_buffer("J_regressor", t(a["J_regressor"]))
        self.register_buffer("lbs_weights", t(a["weights"]))
        parents = a["parents"] if "parents" in a else np.asarray(a["kintree_table"])[0].astype(np.int64)
        parents = np.asarray(parents).astype(np.int64).copy()
        parents[0] = -1
        self.register_buffer("parents", torch.as_tensor(parents, dtype=torch.long))
        self.register_buffer("faces_tensor", torch.as_tensor(np.asarray(a["f"]).astype(np.int64)))
        comps = t(a["hands_components"])[:num_pca_comps]
        self.register_buffer("hand_components", comps)
        mean = torch.zeros(45, dtype=dtype) if flat_hand_mean else t(a["hands_mean"])
        self.register_buffer("hand_mean", mean)
        self.register_buffer("pose_mean", torch.cat([torch.zeros(3, dtype=dtype), mean]))

    def forward(self, betas=None, global_orient=None, hand_pose=None, transl=None, **kwargs):
        if self.use_pca:
            hand_pose = torch.einsum("bi,ij->bj", [hand_pose, self.hand_components])
        full_pose = torch.cat([global_orient, hand_pose], dim=1) + self.pose_mean
        vertices, joints = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs,
                               self.J_regressor, self.parents, self.lbs_weights)
        if transl is not None:
            joints = joints + transl.unsqueeze(dim=1)
            vertices = vertices + transl.unsqueeze(dim=1)
        return vertices, joints, betas, transl, global_orient, full_pose


def load(model_path=None, is_right=True, model_type="mano", num_pca_comps=6, use_pca=True, batch_size=1,
         flat_hand_mean=False, **kwargs):
    return ManoLayer(model_path, num_pca_comps=num_pca_comps, use_pca=use_pca, flat_hand_mean=flat_hand_mean,
                     is_right=is_right, batch_size=batch_size)
