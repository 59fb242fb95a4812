"""ORACLE (test infrastructure; never imported by the product).

Self-contained torch-CPU restatement of the reference's per-iteration hot path so that it can
travel to the GPU box (where /root/reference does not exist):

    optimize_hand_object loop      /root/reference/homan/jointopt.py:128-192
    HOMan.forward                  /root/reference/homan/homan.py:421-508
    get_verts_object / _hand       /root/reference/homan/homan.py:298-307,341-382
    rot6d_to_matrix                /root/reference/homan/utils/geometry.py:9-27
    compute_transformation_persp   /root/reference/homan/utils/camera.py:108-139
    Losses.*                       /root/reference/homan/losses.py:98-242
    lossutils.*                    /root/reference/homan/lossutils.py:18-64,107-130
    SDFSceneLoss.forward           /root/reference/homan/interactions/scenesdf.py:77-148
    contactloss.compute_contact_loss (default-argument path)
                                   /root/reference/homan/interactions/contactloss.py:149-309

It is checked against the UNMODIFIED reference Python in this container
(tests/test_oracle_vs_reference.py, golden vectors under tests/golden/).  Extension over the
reference: a leading problem axis P (clips x inits); every normaliser is evaluated per problem, so
problem p reproduces what the reference computes when given problem p alone.  One hand per frame
(H = 1), `hand_proj_mode="persp"`, `optimize_mano=True`, `optimize_mano_beta=True`,
`optimize_object_scale=False` (the README configuration, README.md:207-238).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import mano_layer, nmr, sdfmod
from .libyana_min import batch_mask_iou, batch_pairwise_dist, batch_proj2d

LOSS_KEYS = ("loss_pca", "loss_smooth_obj", "loss_smooth_hand", "loss_collision", "loss_contact",
             "loss_v2d_hand", "loss_sil_obj", "loss_sil_hand", "loss_inter", "loss_scale_obj",
             "loss_scale_hand")
REND_SIZE = 256          # /root/reference/homan/constants.py:32
INTER_EXPANSION = 0.2    # /root/reference/homan/losses.py:93
INTER_Z_THRESH = 3.0     # /root/reference/homan/losses.py:88
SDF_GRID = 32            # /root/reference/homan/interactions/scenesdf.py:14
SDF_SCALE_FACTOR = 0.2   # /root/reference/homan/interactions/scenesdf.py:77
COLLISION_THRESH = 0.020  # /root/reference/homan/interactions/contactloss.py:156


def rot6d_to_matrix(rot_6d):
    r = rot_6d.view(-1, 3, 2)
    a1, a2 = r[:, :, 0], r[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=-1)  # intended dim=-1 (reference relies on the legacy default)
    return torch.stack((b1, b2, b3), dim=-1)


def transform_persp(meshes, translations, rotations, scale):
    """(s * v) @ R + t and its mesh-detached twin (camera.py:108-139)."""
    scaled = scale.view(-1, 1, 1) * meshes
    out = torch.matmul(scaled, rotations) + translations
    out_det = torch.matmul(scaled.detach().clone(), rotations) + translations
    return out, out_det


class ManoPca:
    """ManoModel.forward_pca with 16 comps (manomodel.py:84-151; homan.py:70)."""

    def __init__(self, asset_right, asset_left=None, ncomps=16):
        self.ncomps = ncomps
        self.layers = {"right": mano_layer.ManoLayer(asset_right, num_pca_comps=ncomps, use_pca=False,
                                                     flat_hand_mean=True)}
        self.means = {"right": torch.as_tensor(np.asarray(asset_right["hands_mean"]), dtype=torch.float32)}
        self.comps = {"right": torch.as_tensor(np.asarray(asset_right["hands_components"]),
                                               dtype=torch.float32)[:ncomps]}
        if asset_left is not None:
            self.layers["left"] = mano_layer.ManoLayer(asset_left, num_pca_comps=ncomps, use_pca=False,
                                                       flat_hand_mean=True)
            self.means["left"] = torch.as_tensor(np.asarray(asset_left["hands_mean"]), dtype=torch.float32)
            self.comps["left"] = torch.as_tensor(np.asarray(asset_left["hands_components"]),
                                                 dtype=torch.float32)[:ncomps]

    def __call__(self, pca_pose, rot, betas, side="right"):
        hand_pose = pca_pose[:, :self.ncomps] @ self.comps[side]
        if side == "left":
            sign = torch.ones(45)
            sign[1::3] = -1
            sign[2::3] = -1
            hand_pose = hand_pose * sign
        hand_pose = hand_pose + self.means[side]
        verts, joints, *_ = self.layers[side](betas=betas, global_orient=rot, hand_pose=hand_pose,
                                              transl=rot.new_zeros(rot.shape[0], 3))
        return verts, joints


def sdf_scene(vertices, faces, scale_factor=SDF_SCALE_FACTOR, grid=SDF_GRID):
    """SDFSceneLoss.forward for a list of objects -> (loss scalar, dist_values dict)."""
    sdf = sdfmod.SDF()
    centers, scales, phis = [], [], []
    with torch.no_grad():
        for v in vertices:
            lo, hi = v.min(1)[0], v.max(1)[0]
            centers.append(((lo + hi) / 2).unsqueeze(1))                       # [T,1,3]
            scales.append(((hi - lo) * ((1 + scale_factor) * 0.5)).max(-1)[0])   # [T]
        for v, f, c, s in zip(vertices, faces, centers, scales):
            local = (v - c) / s.view(-1, 1, 1)
            phis.append(sdf(f, local.contiguous(), grid).clamp(0))
    loss = vertices[0].new_zeros(())
    dist_values = {}
    n = len(vertices)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            local = (vertices[j] - centers[i]) / scales[i].view(-1, 1, 1)
            d = F.grid_sample(phis[i].unsqueeze(1), local.view(local.shape[0], local.shape[1], 1, 1, 3),
                              align_corners=False)
            dist_values[(i, j)] = d[:, 0, :, 0, 0] * scales[i].unsqueeze(1)
            loss = loss + d.sum()
    return loss, dist_values


def project_bbox(verts, K, expansion):
    """losses.py:20-49 with R = I, t = 0, zero distortion, orig_size = 1."""
    world = verts * verts.new_tensor([[[1.0, -1.0, 1.0]]])
    proj = nmr.projection(world, K, torch.eye(3)[None], torch.zeros(1, 3), torch.zeros(1, 5), 1)[:, :, :2]
    boxes = torch.cat([proj.min(1)[0], proj.max(1)[0]], 1)
    center = (boxes[:, :2] + boxes[:, 2:]) / 2
    extent = (boxes[:, 2:] - boxes[:, :2]) / 2 * (1 + expansion)
    return torch.cat([center - extent, center + extent], 1)


def _iou_xyxy(b1, b2):
    a1 = (b1[2] - b1[0]) * (b1[3] - b1[1])
    a2 = (b2[2] - b2[0]) * (b2[3] - b2[1])
    lt = torch.max(b1[:2], b2[:2])
    rb = torch.min(b1[2:], b2[2:])
    wh = (rb - lt).clamp_min(0)
    inter = wh[0] * wh[1]
    return inter / (a1 + a2 - inter)


def _dist_z(v1, v2):
    a, b, c, d = v1[:, 2].min(), v1[:, 2].max(), v2[:, 2].min(), v2[:, 2].max()
    if d >= a and b >= c:
        return 0.0
    return torch.min(torch.abs(c - b), torch.abs(a - d))


class ClipModel:
    """One clip (T frames, one hand, one object): parameters + forward, mirroring HOMan."""

    def __init__(self, prob, mano, closed_faces):
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32).clone()  # noqa: E731
        self.T = T = prob["obj_t"].shape[0]
        self.side = prob.get("side", "right")
        self.mano = mano
        self.image_size = float(prob.get("image_size", 640))
        # parameters (names as in homan.py:66-153)
        self.translations_object = t(prob["obj_t"]).view(T, 1, 3).requires_grad_()
        self.rotations_object = t(prob["obj_R"])[:, :, :2].contiguous().requires_grad_()
        self.translations_hand = t(prob["hand_t"]).view(T, 1, 3).requires_grad_()
        self.rotations_hand = t(prob["hand_R"])[:, :, :2].contiguous().requires_grad_()
        self.mano_pca_pose = t(prob["pca"]).requires_grad_()
        self.mano_rot = t(prob["mano_rot"]).requires_grad_()
        self.mano_trans = t(prob["mano_trans"]).requires_grad_()
        self.mano_betas = torch.zeros(T, 10).requires_grad_()  # re-zeroed: homan.py:108
        self.int_scales_object = torch.ones(1)
        self.int_scales_hand = torch.ones(1)
        # constants
        self.verts_object_og = t(prob["obj_verts_can"]).unsqueeze(0).repeat(T, 1, 1)
        self.faces_object = torch.as_tensor(np.asarray(prob["obj_faces"]).astype(np.int32)).unsqueeze(0).repeat(T, 1, 1)
        self.faces_hand = torch.as_tensor(np.asarray(prob["hand_faces"]).astype(np.int32)).unsqueeze(0)
        self.closed_faces = torch.as_tensor(np.asarray(closed_faces).astype(np.int32))
        tm_o, tm_h = t(prob["target_masks_object"]), t(prob["target_masks_hand"])
        self.ref_mask_object, self.keep_mask_object = (tm_o > 0).float(), (tm_o >= 0).float()
        self.ref_mask_hand, self.keep_mask_hand = (tm_h > 0).float(), (tm_h >= 0).float()
        self.camintr_rois_object = t(prob["K_roi_obj"])
        self.camintr_rois_hand = t(prob["K_roi_hand"])
        self.camintr = t(prob["camintr"])
        self.ref_verts2d_hand = t(prob["verts2d"])
        self.renderer = nmr.Renderer(image_size=REND_SIZE, K=self.camintr.clone(), R=torch.eye(3)[None],
                                     t=torch.zeros(1, 3), orig_size=1)

    def named_parameters(self):
        return [(k, getattr(self, k)) for k in
                ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
                 "mano_pca_pose", "mano_rot", "mano_trans", "mano_betas")]

    def verts_object(self):
        return transform_persp(self.verts_object_og, self.translations_object,
                               rot6d_to_matrix(self.rotations_object), self.int_scales_object.abs())

    def verts_hand(self):
        v, _ = self.mano(self.mano_pca_pose, self.mano_rot, self.mano_betas, self.side)
        v = v + self.mano_trans.unsqueeze(1)
        return transform_persp(v, self.translations_hand, rot6d_to_matrix(self.rotations_hand),
                               self.int_scales_hand)

    def forward(self, lw):
        T = self.T
        losses, metrics = {}, {}
        verts_object, _ = self.verts_object()
        verts_hand, verts_hand_det = self.verts_hand()
        verts_hand_ds = verts_hand  # scale is a buffer: detach_scale changes nothing (homan.py:359-362)
        if lw["lw_pca"] > 0:
            losses["loss_pca"] = (self.mano_pca_pose ** 2).mean()
        if lw["lw_smooth_hand"] > 0 or lw["lw_smooth_obj"] > 0:
            losses["loss_smooth_obj"] = ((verts_object[1:] - verts_object[:-1]) ** 2).mean()
            losses["loss_smooth_hand"] = ((verts_hand[1:] - verts_hand[:-1]) ** 2).mean()
        if lw["lw_collision"] > 0:
            sdf_loss, _ = sdf_scene([verts_hand_ds, verts_object.detach()], [self.closed_faces, self.faces_object[0]])
            losses["loss_collision"] = sdf_loss.mean()
        if lw["lw_contact"] > 0:
            losses["loss_contact"] = self.contact(verts_hand_ds, verts_object)
        if lw["lw_v2d_hand"] > 0:
            proj = batch_proj2d(verts_hand, self.camintr)
            tar = self.ref_verts2d_hand / self.image_size
            losses["loss_v2d_hand"] = ((proj - tar) ** 2).sum(-1).mean()
            metrics["v2d_hand"] = (proj * self.image_size - self.ref_verts2d_hand).norm(2, -1).mean().item()
        if lw["lw_sil_obj"] > 0:
            rend = self.renderer(verts_object, self.faces_object, K=self.camintr_rois_object, mode="silhouettes")
            image = self.keep_mask_object * rend
            losses["loss_sil_obj"] = ((image - self.ref_mask_object) ** 2).sum() / self.keep_mask_object.sum() / T
            metrics["iou_object"] = batch_mask_iou(image, self.ref_mask_object).mean().item()
        if lw.get("lw_sil_hand", 0) > 0:
            # intended semantics of the (unused, buggy) compute_sil_loss_hand, losses.py:166-181
            rend = self.renderer(verts_hand, self.faces_hand.repeat(T, 1, 1), K=self.camintr_rois_hand,
                                 mode="silhouettes")
            image = self.keep_mask_hand * rend
            per = ((image - self.ref_mask_hand) ** 2).sum((1, 2)) / self.keep_mask_hand.sum((1, 2))
            losses["loss_sil_hand"] = per.sum() / T
            metrics["iou_hand"] = batch_mask_iou(image, self.ref_mask_hand).mean().item()
        if lw["lw_inter"] > 0:
            l, m = self.interaction(verts_hand_det, verts_object.detach())
            losses["loss_inter"] = l
            metrics["handobj_maxdist"] = m
        if lw["lw_scale_obj"] > 0:
            losses["loss_scale_obj"] = ((self.int_scales_object - 1.0) ** 2).sum() / 1
        if lw["lw_scale_hand"] > 0:
            losses["loss_scale_hand"] = ((self.int_scales_hand - 1.0) ** 2).sum() / 1
        return losses, metrics

    def contact(self, hand, obj):
        """Default path of compute_contact_loss: phi is clamped >= 0 so `exterior` is always False;
        loss = mean over (T, 778) of 0.02 * tanh(|nearest object vertex - hand vertex| / 0.02)."""
        d = batch_pairwise_dist(hand, obj)
        idx = d.min(2)[1]
        close = torch.gather(obj, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
        anchor = torch.norm(close - hand, 2, 2)
        vals = COLLISION_THRESH * torch.tanh(anchor / COLLISION_THRESH)
        return vals.sum() / vals.numel()

    def interaction(self, hand_det, obj):
        K = self.renderer.K
        with torch.no_grad():
            bo = project_bbox(obj, K, INTER_EXPANSION)
            bh = project_bbox(hand_det, K, INTER_EXPANSION)
            flags = []
            for t in range(self.T):
                iou = _iou_xyxy(bo[t], bh[t])
                zd = _dist_z(obj[t], hand_det[t])
                flags.append(bool((iou > 0) and (zd < INTER_Z_THRESH)))
        loss = hand_det.new_zeros(1)
        for t, fl in enumerate(flags):
            if fl:
                loss = loss + F.mse_loss(hand_det[t].mean(0), obj[t].mean(0))
        with torch.no_grad():
            md = torch.sqrt(batch_pairwise_dist(hand_det, obj)).min(1)[0].min(1)[0]
        return loss, md.max().item()


def make_optimizer(model, lr):
    """Three Adam groups of jointopt.py:128-151 (mano_rot / mano_trans match no group)."""
    named = model.named_parameters()
    rigid = [v for k, v in named if "mano" not in k and "rotation" not in k]
    rots = [v for k, v in named if "rotation" in k and "mano" not in k]
    return torch.optim.Adam([{"params": rigid, "lr": lr},
                             {"params": [model.mano_pca_pose, model.mano_betas], "lr": lr * 10},
                             {"params": rots, "lr": lr * 10}])


def problem_slice(batch, p):
    """Problem p of a batched problem dict (see homan_b200.problem.make_batch)."""
    out = {k: batch[k] for k in ("obj_verts_can", "obj_faces", "hand_faces", "side", "image_size") if k in batch}
    if np.asarray(batch["obj_verts_can"]).ndim == 3:   # one object per clip (clip-major multi-clip batch)
        c = int(np.asarray(batch["clip_of_problem"])[p])
        out["obj_verts_can"], out["obj_faces"] = batch["obj_verts_can"][c], batch["obj_faces"][c]
    for k in ("obj_t", "obj_R", "hand_t", "hand_R", "pca", "mano_rot", "mano_trans", "betas",
              "target_masks_object", "target_masks_hand", "K_roi_obj", "K_roi_hand", "camintr", "verts2d"):
        out[k] = batch[k][p]
    return out


def fit(batch, loss_weights, num_iterations, lr=1e-2, mano_assets=None, problems=None, record_grads=False):
    """Runs the reference loop independently on every problem of `batch`.
    Returns {"losses": {key: [iters, P]}, "metrics": {...}, "total": [iters, P], "params": {name: [P, ...]}}."""
    torch.manual_seed(0)
    assets = mano_assets or {"right": batch["mano_asset"]}
    mano = ManoPca(assets["right"], assets.get("left"))
    closed = assets[batch.get("side", "right")]["closed_faces"]
    P = batch["obj_t"].shape[0]
    problems = list(range(P)) if problems is None else problems
    hist = {"losses": {}, "metrics": {}, "total": np.zeros((num_iterations, len(problems)), np.float64), "params": {},
            "grads0": {}}
    for col, p in enumerate(problems):
        model = ClipModel(problem_slice(batch, p), mano, closed)
        opt = make_optimizer(model, lr)
        for it in range(num_iterations):
            opt.zero_grad()
            losses, metrics = model.forward(loss_weights)
            total = sum(v * loss_weights[k.replace("loss", "lw")] for k, v in losses.items())
            for k, v in losses.items():
                hist["losses"].setdefault(k, np.zeros((num_iterations, len(problems))))[it, col] = float(v)
            for k, v in metrics.items():
                hist["metrics"].setdefault(k, np.zeros((num_iterations, len(problems))))[it, col] = float(v)
            hist["total"][it, col] = float(total)
            total.backward()
            if record_grads and it == 0:
                for k, v in model.named_parameters():
                    g = v.grad if v.grad is not None else torch.zeros_like(v)
                    hist["grads0"].setdefault(k, []).append(g.detach().numpy().copy())
            opt.step()
        for k, v in model.named_parameters():
            hist["params"].setdefault(k, []).append(v.detach().numpy().copy())
    hist["params"] = {k: np.stack(v) for k, v in hist["params"].items()}
    hist["grads0"] = {k: np.stack(v) for k, v in hist["grads0"].items()}
    return hist


def evaluate(batch, loss_weights, mano_assets=None, problems=None):
    """One forward + backward at the initial parameters: per-problem losses and parameter gradients."""
    return fit(batch, loss_weights, 1, lr=0.0, mano_assets=mano_assets, problems=problems, record_grads=True)
