"""Generates tests/golden/ref_pose_init.npz by running the UNMODIFIED reference
/root/reference/homan/pose_optimization.py::find_optimal_pose on CPU under oracle/refshim.py
(build container only). The fixture holds the complete inputs and, per iteration, the per-candidate loss terms
and IoU the reference evaluated, plus the parameters it returned.

    python scripts/make_golden_pose.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from homan_b200 import synth  # noqa: E402
from oracle import libyana_min, nmr, refshim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_INITS, N_ITERS, SEED = 6, 6, 4000


def make_case():
    rng = np.random.default_rng(SEED)
    verts, faces = synth.make_object("ellipsoid80")
    image_size, focal = 640, 600.0
    K = np.array([[focal, 0, 320.0], [0, focal, 320.0], [0, 0, 1]], dtype=np.float32)
    R_gt = synth._random_rotation(rng).astype(np.float32)
    t_gt = np.array([0.03, -0.02, 0.55], dtype=np.float32)
    v_gt = verts.astype(np.float32) @ R_gt + t_gt
    uv = synth.project_np(v_gt.astype(np.float64), K.astype(np.float64))
    lo, hi = uv.min(0), uv.max(0)
    bbox = np.array([lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]], dtype=np.float32)  # xywh
    x, y, b = synth._square_roi(uv)
    square_bbox = np.array([x, y, b, b], dtype=np.float32)
    K_roi = libyana_min.get_K_crop_resize(torch.from_numpy(K)[None], torch.tensor([[x, y, x + b, y + b]]), [256])
    K_roi[:, :2] /= 256
    rend = nmr.Renderer(image_size=256, K=K_roi, R=torch.eye(3)[None], t=torch.zeros(1, 3), orig_size=1,
                        anti_aliasing=False)
    mask = rend(torch.from_numpy(v_gt)[None], torch.from_numpy(faces.astype(np.int32))[None], mode="silhouettes")[0].numpy()
    mask = (mask > 0.5).astype(np.float32)
    mask[:, 100:118] = -1  # occluded band
    # candidate rotations: perturbations of the ground truth, one far off (partly off-screen after translation init)
    rots = []
    for i in range(N_INITS):
        dr = synth._axis_angle_to_mat(rng.normal(size=3) * (0.15 if i else 0.02))[0]
        rots.append((R_gt.astype(np.float64) @ dr).astype(np.float32))
    return dict(vertices=verts.astype(np.float32), faces=faces.astype(np.int64), mask=mask, bbox=bbox,
                square_bbox=square_bbox, image_size=np.array([image_size, image_size]), K=K,
                rotations_init=np.stack(rots))


def main():
    assert refshim.reference_available(), "needs /root/reference"
    assets = {"right": synth.make_mano_asset(0, "right"), "left": synth.make_mano_asset(1, "left")}
    scratch = tempfile.mkdtemp(prefix="homan_golden_pose_")
    refshim.install(scratch, assets)
    import homan.pose_optimization as po
    case = make_case()
    rec = {"mask": [], "offscreen": [], "chamfer": [], "iou": [], "rot": [], "trans": []}
    orig_forward = po.PoseOptimizer.forward

    def forward(model):  # observes, does not alter
        loss_dict, iou, image = orig_forward(model)
        for k in ("mask", "offscreen", "chamfer"):
            rec[k].append(loss_dict[k].detach().numpy().copy())
        rec["iou"].append(iou.detach().numpy().copy())
        rec["rot"].append(model.rotations.detach().numpy().copy())
        rec["trans"].append(model.translations.detach().numpy().copy())
        return loss_dict, iou, image

    po.PoseOptimizer.forward = forward
    out = {}
    for sort_best in (False, True):
        for v in rec.values():
            v.clear()
        model = po.find_optimal_pose(
            vertices=torch.from_numpy(case["vertices"]), faces=torch.from_numpy(case["faces"]), mask=case["mask"],
            bbox=case["bbox"], square_bbox=case["square_bbox"], image_size=tuple(case["image_size"]), K=case["K"],
            num_iterations=N_ITERS, num_initializations=N_INITS, debug=False, viz=False,
            viz_folder=os.path.join(scratch, "viz"), sort_best=sort_best,
            rotations_init=torch.from_numpy(case["rotations_init"]))
        tag = "sorted" if sort_best else "plain"
        out[f"{tag}_rotations"] = model.rotations.detach().numpy()
        out[f"{tag}_translations"] = model.translations.detach().numpy()
        if not sort_best:
            for k, v in rec.items():
                out["it_" + k] = np.stack(v)  # [iters, N, ...]: what the reference evaluated at each iteration
            out["K_roi"] = model.K.detach().numpy()
    po.PoseOptimizer.forward = orig_forward
    # PoseOptimizer called directly with candidates pushed partly off-screen / behind the camera: loss terms and
    # the gradients of one backward (exercises compute_offscreen_loss, pose_optimization.py:112-134)
    K_roi = torch.from_numpy(out["K_roi"])
    rot6 = torch.from_numpy(case["rotations_init"][:4, :, :2].copy())
    trans = torch.from_numpy(out["it_trans"][0][:4].copy())
    trans[1, 0, 0] += 0.07     # sticks out on the right
    trans[2, 0, 1] -= 0.09     # sticks out at the top
    trans[3, 0, 2] = -0.2      # behind the camera
    model = po.PoseOptimizer(ref_image=case["mask"], vertices=torch.from_numpy(case["vertices"]),
                             faces=torch.from_numpy(case["faces"]), textures=torch.ones(80, 1, 1, 1, 3),
                             rotation_init=rot6, translation_init=trans, num_initializations=4, K=K_roi)
    loss_dict, iou, image = model()
    sum(loss_dict.values()).sum().backward()
    out.update({"po_rot6d": rot6.numpy(), "po_trans": trans.numpy(), "po_mask": loss_dict["mask"].detach().numpy(),
                "po_offscreen": loss_dict["offscreen"].detach().numpy(), "po_iou": iou.numpy(),
                "po_grad_rot": model.rotations.grad.numpy(), "po_grad_trans": model.translations.grad.numpy(),
                "po_image_sum": image.detach().sum((1, 2)).numpy()})
    out.update({"in_" + k: v for k, v in case.items()})
    out["iters"] = np.array(N_ITERS)
    path = os.path.join(GOLDEN, "ref_pose_init.npz")
    np.savez_compressed(path, **out)
    print(path, {k: np.asarray(v).shape for k, v in out.items()})
    print("mask loss per iteration (candidate 0..):", out["it_mask"][:, :3])
    print("offscreen:", out["it_offscreen"].max(), out["po_offscreen"], out["po_mask"])


def main_chamfer():
    """tests/golden/ref_pose_chamfer.npz: PoseOptimizer with the max-pool edge / EDT chamfer term switched on
    (pose_optimization.py:84-88,136-149; lw_chamfer is 0 in every call of the reference itself), on the inputs of
    ref_pose_init.npz: loss terms and the gradients of one backward."""
    assert refshim.reference_available(), "needs /root/reference"
    assets = {"right": synth.make_mano_asset(0, "right"), "left": synth.make_mano_asset(1, "left")}
    refshim.install(tempfile.mkdtemp(prefix="homan_golden_pose_"), assets)
    import homan.pose_optimization as po
    g = np.load(os.path.join(GOLDEN, "ref_pose_init.npz"))
    rot6 = torch.from_numpy(g["po_rot6d"].copy())
    trans = torch.from_numpy(g["it_trans"][0][:4].copy())
    trans[1, 0, 0] += 0.02     # candidates around the target: edges near, but not on, the reference's
    trans[2, 0, 1] -= 0.03
    trans[3, 0, 2] += 0.05
    out = {"po_rot6d": rot6.numpy(), "po_trans": trans.numpy(), "lw_chamfer": np.float32(0.5), "kernel_size": np.int64(7),
           "power": np.float32(0.25)}
    model = po.PoseOptimizer(ref_image=g["in_mask"], vertices=torch.from_numpy(g["in_vertices"]),
                             faces=torch.from_numpy(g["in_faces"]), textures=torch.ones(80, 1, 1, 1, 3),
                             rotation_init=rot6, translation_init=trans, num_initializations=4,
                             K=torch.from_numpy(g["K_roi"]), kernel_size=7, power=0.25, lw_chamfer=0.5)
    loss_dict, iou, image = model()
    loss_dict["chamfer"].sum().backward()
    out.update({"po_chamfer": loss_dict["chamfer"].detach().numpy(), "po_mask": loss_dict["mask"].detach().numpy(),
                "po_offscreen": loss_dict["offscreen"].detach().numpy(), "po_iou": iou.numpy(),
                "po_grad_rot_chamfer": model.rotations.grad.numpy().copy(),
                "po_grad_trans_chamfer": model.translations.grad.numpy().copy(),
                "edt_ref_edge": model.edt_ref_edge[0].numpy()})
    path = os.path.join(GOLDEN, "ref_pose_chamfer.npz")
    np.savez_compressed(path, **out)
    print(path, {k: np.asarray(v).shape for k, v in out.items()})
    print("chamfer:", out["po_chamfer"], "grad trans:", out["po_grad_trans_chamfer"].reshape(4, 3))


if __name__ == "__main__":
    if "--chamfer" in sys.argv:
        main_chamfer()
    else:
        main()
