"""ORACLE (test infrastructure; never imported by the product).

CPU restatement (torch + oracle/nmr.py) of the object-pose multi-init fitter of the reference,
/root/reference/homan/pose_optimization.py: PoseOptimizer.forward (:136-149), compute_offscreen_loss (:112-134),
apply_transformation (:105-110) and the optimisation loop of find_optimal_pose (:332-356, Adam lr 1e-2 over
rotations [N,3,2] and translations [N,1,3], loss summed over the candidates, best-ever candidate read AFTER
optimizer.step()). The chamfer term is omitted (lw_chamfer = 0 in every reference call). Pinned by
tests/golden/ref_pose_init.npz, recorded from the unmodified reference (scripts/make_golden_pose.py).
"""
import numpy as np
import torch

from . import nmr


def rot6d_to_matrix(r6):
    """/root/reference/homan/utils/geometry.py:9-27."""
    r6 = r6.view(-1, 3, 2)
    a1, a2 = r6[:, :, 0], r6[:, :, 1]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=2)


def forward(vertices, faces, ref_image, K_roi, rotations, translations, far=100.0):
    """-> (loss_dict {mask, offscreen} [N], iou [N], image [N,R,R]); differentiable w.r.t. rotations / translations."""
    N = rotations.shape[0]
    R = ref_image.shape[0]
    image_ref = torch.from_numpy((np.asarray(ref_image) > 0).astype(np.float32))[None].repeat(N, 1, 1)
    keep = torch.from_numpy((np.asarray(ref_image) >= 0).astype(np.float32))[None].repeat(N, 1, 1)
    verts = torch.matmul(vertices[None].repeat(N, 1, 1), rot6d_to_matrix(rotations)) + translations
    renderer = nmr.Renderer(image_size=R, K=K_roi, R=torch.eye(3)[None], t=torch.zeros(1, 3), orig_size=1,
                            anti_aliasing=False)
    image = keep * renderer(verts, faces[None].repeat(N, 1, 1), mode="silhouettes")
    mask = torch.sum((image - image_ref) ** 2, dim=(1, 2))
    a = image.detach()
    iou = (a * image_ref).sum((1, 2)) / ((a + image_ref).clamp(0, 1).sum((1, 2)) + 1e-6)
    proj = nmr.projection(verts, K_roi, torch.eye(3)[None], torch.zeros(1, 3), torch.zeros(1, 5), 1)
    xy, z = proj[:, :, :2], proj[:, :, 2:]
    zeros = torch.zeros_like(z)
    off = (torch.max(xy - 1, zeros).sum(dim=(1, 2)) + torch.max(-1 - xy, zeros).sum(dim=(1, 2)) +
           torch.max(-z, zeros).sum(dim=(1, 2)) + torch.max(z - far, zeros).sum(dim=(1, 2)))
    return {"mask": mask, "offscreen": 100000 * off}, iou, image


def fit(vertices, faces, ref_image, K_roi, rot6d_init, trans_init, num_iterations, lr=1e-2):
    """The loop of find_optimal_pose. Returns per-iteration totals [iters,N], final parameters, best-ever."""
    vertices = torch.as_tensor(vertices).float()
    faces = torch.as_tensor(np.asarray(faces)).int()
    K_roi = torch.as_tensor(K_roi).float().view(1, 3, 3)
    rot = torch.as_tensor(rot6d_init).float().clone().requires_grad_()
    tr = torch.as_tensor(trans_init).float().clone().view(-1, 1, 3).requires_grad_()
    opt = torch.optim.Adam([rot, tr], lr=lr)
    hist, ious = [], []
    best = {"loss": np.inf, "rot": None, "trans": None}
    for _ in range(num_iterations):
        opt.zero_grad()
        ld, iou, _ = forward(vertices, faces, ref_image, K_roi, rot, tr)
        losses = ld["mask"] + ld["offscreen"]
        losses.sum().backward()
        opt.step()
        if losses.min().item() < best["loss"]:
            ind = int(torch.argmin(losses))
            best = {"loss": losses[ind].item(), "rot": rot[ind].detach().clone().numpy(),
                    "trans": tr[ind].detach().clone().numpy()}
        hist.append(losses.detach().numpy().copy())
        ious.append(iou.numpy().copy())
    return {"total": np.stack(hist), "iou": np.stack(ious), "rotations": rot.detach().numpy(),
            "translations": tr.detach().numpy(), "best": best}
