"""Object-pose multi-init fitter: the reference's per-frame initialiser of the object pose
(/root/reference/homan/pose_optimization.py) on the homan_b200 kernels.

Same public surface as the reference module:

  PoseOptimizer(nn.Module)      pose_optimization.py:37-160   forward() -> (loss_dict, iou, image)
  find_optimal_pose(...)        pose_optimization.py:219-383  N random inits x num_iterations of Adam
  find_optimal_poses(...)       pose_optimization.py:386-488  frame by frame, best motion over the clip

plus the helpers it needs (`compute_random_rotations`, homan/utils/geometry.py:89-134;
`TCO_init_from_boxes_zup_autodepth`, homan/lib3d/optitrans.py:83-127; `get_K_crop_resize` of libyana).
`find_optimal_pose` runs on `PoseFitEngine`: one iteration = rigid placement -> projection -> hard z-buffered
silhouette (anti-aliasing off) -> masked L2 + off-screen penalty -> backward -> Adam -> best-ever tracking, a
fixed sequence of C-ABI kernels replayed as a CUDA graph, all N candidates in one batch. GPU only, no fallback.
The chamfer term (max-pool edges x distance transform, lw_chamfer = 0 in every reference call) is evaluated by the
drop-in module only (torch max-pool on the rasteriser's output); the fused fitter runs the reference's configuration.
Not provided: the debug plots.
"""
import math

import numpy as np
import torch
from torch import nn

from . import _lib, ops
from ._lib import call, current_stream, ptr
from .engine import NPART, PART, REND_SIZE

PART_OFFSCREEN = 14          # include/homan_b200.h HM_PART_OFFSCREEN
OFFSCREEN_WEIGHT = 100000.0  # pose_optimization.py:148


# ------------------------------------------------------------------------------------------ helpers
def rot6d_to_matrix(rot_6d):
    """homan/utils/geometry.py:9-27 ([B,3,2] -> [B,3,3], columns b1 b2 b3)."""
    rot_6d = rot_6d.view(-1, 3, 2)
    a1, a2 = rot_6d[:, :, 0], rot_6d[:, :, 1]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(1, keepdim=True) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=2)


def matrix_to_rot6d(rotmat):
    """homan/utils/geometry.py:30-40."""
    return rotmat.view(-1, 3, 3)[:, :, :2]


def compute_random_rotations(B=10, upright=False, generator=None, device="cuda"):
    """homan/utils/geometry.py:89-134, uniform branch (J. Arvo, "Fast Random Rotation Matrices")."""
    if upright:
        # (upstream's upright branch calls `euler_angles_to_matrix`, which homan/utils/geometry.py neither defines nor
        #  imports: it raises NameError there; find_optimal_pose never asks for it)
        raise NotImplementedError("compute_random_rotations: upright=True (broken upstream, unused by find_optimal_pose)")
    x1, x2, x3 = torch.split(torch.rand(3 * B, generator=generator).to(device), B)
    tau = 2 * math.pi
    zeros, ones = torch.zeros_like(x1), torch.ones_like(x1)
    R = torch.stack((torch.stack((torch.cos(tau * x1), torch.sin(tau * x1), zeros), 1),
                     torch.stack((-torch.sin(tau * x1), torch.cos(tau * x1), zeros), 1),
                     torch.stack((zeros, zeros, ones), 1)), 1)
    v = torch.stack((torch.cos(tau * x2) * torch.sqrt(x3), torch.sin(tau * x2) * torch.sqrt(x3),
                     torch.sqrt(1 - x3)), 1)
    H = torch.eye(3, device=x1.device).repeat(B, 1, 1) - 2 * v.unsqueeze(2) * v.unsqueeze(1)
    return -torch.matmul(H, R)


def batch_proj2d(verts, camintr):
    hom = camintr.bmm(verts.transpose(1, 2)).transpose(1, 2)
    return hom[:, :, :2] / hom[:, :, 2:]


def TCO_init_from_boxes_zup_autodepth(boxes_2d, model_points_3d, K):
    """homan/lib3d/optitrans.py:83-127: translation that makes the projected bbox of the (rotated) model match the
    detected box (xywh, pixels), 10 fixed-point iterations. Returns [B,3]."""
    pts = torch.as_tensor(model_points_3d).float()
    bsz, dev = pts.shape[0], pts.device
    K = torch.as_tensor(np.asarray(K) if not torch.is_tensor(K) else K).float().to(dev)
    boxes = torch.as_tensor(np.asarray(boxes_2d) if not torch.is_tensor(boxes_2d) else boxes_2d).float().to(dev)
    if boxes.dim() == 1:
        boxes = boxes.unsqueeze(0)
    if boxes.shape[0] != bsz:
        boxes = boxes.repeat(bsz, 1)
    if K.dim() == 2:
        K = K.unsqueeze(0)
    if K.shape[0] != bsz:
        K = K.repeat(bsz, 1, 1)
    boxes = torch.stack([boxes[:, 0], boxes[:, 1], boxes[:, 0] + boxes[:, 2], boxes[:, 1] + boxes[:, 3]], 1)
    diag_bb = (boxes[:, [2, 3]] - boxes[:, [0, 1]]).norm(2, -1)
    centers = (boxes[:, [0, 1]] + boxes[:, [2, 3]]) / 2
    fxfy, cxcy = K[:, [0, 1], [0, 1]], K[:, [0, 1], [2, 2]]
    z = fxfy.new_ones(bsz, 1)
    xy = ((centers - cxcy) * z) / fxfy
    trans = torch.cat([xy, z], 1)
    for _ in range(10):
        proj = batch_proj2d(pts + trans.unsqueeze(1), K)
        lo, hi = proj.min(1)[0], proj.max(1)[0]
        diag_proj = (lo - hi).norm(2, -1)
        z = z + z * (diag_proj / diag_bb - 1).unsqueeze(-1)
        xy = xy + ((centers - (lo + hi) / 2) * z) / fxfy
        trans = torch.cat([xy, z], 1)
    return trans


def get_K_crop_resize(K, boxes, crop_resize):
    """libyana.lib3d.kcrop.get_K_crop_resize (un-vendored; recalled): intrinsics of the
    crop `boxes` (xyxy) resized to crop_resize. K [B,3,3], boxes [B,4] -> [B,3,3]."""
    K, boxes = K.float(), boxes.float()
    new_K = K.clone()
    final_w, final_h = float(max(crop_resize)), float(min(crop_resize))
    crop_w, crop_h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    crop_cj, crop_ci = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_w - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_h - 1) / 2 - crop_ci
    sx, sy = final_w / crop_w, final_h / crop_h
    new_K[:, 0, 0] = sx * K[:, 0, 0]
    new_K[:, 1, 1] = sy * K[:, 1, 1]
    new_K[:, 0, 2] = (final_w - 1) / 2 + sx * (cx - (crop_w - 1) / 2)
    new_K[:, 1, 2] = (final_h - 1) / 2 + sy * (cy - (crop_h - 1) / 2)
    return new_K


def _mask_to_int8(ref_image):
    m = np.asarray(ref_image)
    out = np.zeros(m.shape, np.int8)
    out[m > 0] = 1
    out[m < 0] = -1
    return out


# ------------------------------------------------------------------------------------------ fused engine
class PoseFitEngine:
    """All N pose candidates of one object against one target mask; step() is one iteration of the loop of
    /root/reference/homan/pose_optimization.py:332-356."""

    def __init__(self, vertices, faces, ref_image, K_roi, rot6d_init, trans_init, lr=1e-2, image_size=None,
                 near=0.1, far=100.0, betas=(0.9, 0.999), eps=1e-8, use_graph=True, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.HomanB200Error("PoseFitEngine needs a CUDA device (there is no CPU path)")
        _lib.lib()
        dev = self.device = torch.device(device)
        f32 = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).float().to(dev).contiguous()  # noqa: E731
        self.mesh = f32(vertices).view(1, -1, 3)
        self.faces = torch.as_tensor(np.asarray(faces) if not torch.is_tensor(faces) else faces).to(dev).int().view(1, -1, 3).contiguous()
        rot6d_init = f32(rot6d_init).view(-1, 3, 2)
        self.N = N = rot6d_init.shape[0]
        trans_init = f32(trans_init).view(-1, 3)
        if trans_init.shape[0] != N:
            trans_init = trans_init.repeat(N, 1)
        self.V, self.F = self.mesh.shape[1], self.faces.shape[1]
        mask = _mask_to_int8(ref_image.detach().cpu().numpy() if torch.is_tensor(ref_image) else ref_image)
        assert mask.shape[0] == mask.shape[1], "Must be square."
        self.R = R = int(image_size or mask.shape[0])
        self.near, self.far = float(near), float(far)
        self.K = f32(K_roi).view(1, 3, 3)
        self.target = torch.from_numpy(mask).to(dev)[None].repeat(N, 1, 1).contiguous()
        self.norm = torch.ones(N, device=dev)
        self.lr, self.adam_betas, self.adam_eps = float(lr), betas, float(eps)

        n = N * 9
        self.flat = torch.empty(n, device=dev)
        self.rotations = self.flat[:N * 6].view(N, 3, 2)
        self.translations = self.flat[N * 6:].view(N, 1, 3)
        self.rotations.copy_(rot6d_init)
        self.translations.copy_(trans_init.view(N, 1, 3))
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.lr_elem = torch.full((n,), self.lr, device=dev)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        sizes = {"grads": n, "g_verts": N * self.V * 3, "g_ndc": N * self.V * 3, "partials": N * NPART}
        self.zero_region = torch.zeros(sum(sizes.values()), device=dev)
        z, off = {}, 0
        for k, m in sizes.items():
            z[k] = self.zero_region[off:off + m]
            off += m
        self.grad_flat, self.g_verts, self.g_ndc, self.partials = z["grads"], z["g_verts"], z["g_ndc"], z["partials"]
        self.grad_rotations = self.grad_flat[:N * 6].view(N, 3, 2)
        self.grad_translations = self.grad_flat[N * 6:].view(N, 1, 3)
        self.verts = torch.empty(N, self.V, 3, device=dev)
        self.ndc = torch.empty(N, self.V, 3, device=dev)
        self.rb = ops.RasterBuffers(N, self.V, self.F, R, False, dev)
        self.ga = torch.empty(N, R, R, device=dev)
        self.losses = torch.zeros(N, NPART, device=dev)
        self.total = torch.zeros(N, device=dev)
        w = torch.zeros(NPART)
        w[PART["sil_obj"]] = 1.0
        w[PART_OFFSCREEN] = OFFSCREEN_WEIGHT
        self.weights_part = w.to(dev)
        self.best = torch.zeros(10, device=dev)
        self.best[0] = float("inf")
        self.best_index = torch.full((1,), -1, dtype=torch.int32, device=dev)
        self.use_graph, self.graph, self.iteration = use_graph, None, 0
        self.gpu_launches_per_step = 0

    def _iteration(self, backward=True):
        s = current_stream()
        N, V = self.N, self.V
        self.zero_region.zero_()
        call("hm_rigid_fwd", ptr(self.mesh), 1, ptr(self.rotations), ptr(self.translations), None, N, V,
             ptr(self.verts), s)
        call("hm_project_fwd", ptr(self.verts), ptr(self.K), 1, None, None, None, 0, 1.0, 1e-9, N, V, ptr(self.ndc), s)
        ops.raster_forward(self.rb, self.ndc, self.faces, True, self.near, self.far)
        pb = self.partials.data_ptr()
        call("hm_sil_loss_fwd_bwd", ptr(self.rb.alpha), ptr(self.target), ptr(self.norm), 1.0, N, self.R,
             pb + 4 * PART["sil_obj"], NPART, pb + 4 * PART["iou_obj"], NPART, ptr(self.ga) if backward else None, s)
        call("hm_offscreen_loss_fwd_bwd", ptr(self.ndc), N, V, self.far, OFFSCREEN_WEIGHT, ptr(self.partials),
             ptr(self.g_ndc) if backward else None, s)
        n = 7
        if backward:
            ops.raster_backward(self.rb, self.ga, self.g_ndc)
            call("hm_project_bwd", ptr(self.verts), ptr(self.K), 1, None, None, 1.0, 1e-9, N, V, ptr(self.g_ndc),
                 ptr(self.g_verts), 1, s)
            call("hm_rigid_bwd", ptr(self.mesh), 1, ptr(self.rotations), None, N, V, ptr(self.g_verts),
                 ptr(self.grad_rotations), ptr(self.grad_translations), s)
            n += 5
        call("hm_finalize_losses", ptr(self.partials), ptr(self.weights_part), N, 1, ptr(self.losses), ptr(self.total),
             ptr(self.step_counter) if backward else None, s)
        n += 1
        if backward:
            call("hm_adam_step", ptr(self.flat), ptr(self.grad_flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 ptr(self.lr_elem), N * 9, self.adam_betas[0], self.adam_betas[1], self.adam_eps,
                 ptr(self.step_counter), s)
            call("hm_track_best", ptr(self.total), N, ptr(self.rotations), ptr(self.translations), ptr(self.best),
                 ptr(self.best_index), s)
            n += 2
        self.gpu_launches_per_step = n
        return n

    def evaluate(self):
        """Forward only at the current parameters (model() of the reference): fills losses / total / alpha."""
        self._iteration(backward=False)
        return self.loss_dict()

    def capture(self):
        saved = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.step_counter, self.best,
                                     self.best_index)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._iteration()
        for dst, src in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.step_counter, self.best,
                             self.best_index), saved):
            dst.copy_(src)
        self.graph = graph

    def step(self):
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._iteration()
        self.iteration += 1

    def loss_dict(self):
        """Per-candidate terms of the last evaluation, in the reference's naming (pose_optimization.py:139-149)."""
        ls = self.losses
        return {"mask": ls[:, PART["sil_obj"]].clone(), "chamfer": torch.zeros_like(ls[:, 0]),
                "offscreen": OFFSCREEN_WEIGHT * ls[:, PART_OFFSCREEN]}

    @property
    def iou(self):
        return self.losses[:, PART["iou_obj"]].clone()

    @property
    def image(self):
        return (self.target >= 0).float() * self.rb.alpha

    def fit(self, num_iterations, record=False):
        hist = torch.zeros(num_iterations, self.N, device=self.device) if record else None
        for it in range(num_iterations):
            self.step()
            if record:
                hist[it].copy_(self.total)
        return hist


# ------------------------------------------------------------------------------------------ reference surface
class _Renderer:
    """The attributes of nr.renderer.Renderer the callers of PoseOptimizer read (model.renderer.K, .far, ...)."""

    def __init__(self, image_size, K):
        self.image_size, self.K, self.anti_aliasing, self.fill_back = image_size, K, False, True
        self.R = torch.eye(3, device=K.device).unsqueeze(0)
        self.t = torch.zeros(1, 3, device=K.device)
        self.dist_coeffs = torch.zeros(1, 5, device=K.device)
        self.orig_size, self.near, self.far = 1, 0.1, 100

    def __call__(self, vertices, faces, textures=None, mode=None, K=None):
        if mode != "silhouettes":
            raise NotImplementedError("homan_b200 PoseOptimizer: silhouettes only (RGB rendering is visualisation)")
        ndc = ops.project(vertices, self.K if K is None else K, self.R, self.t, self.dist_coeffs, self.orig_size)
        return ops.rasterize_silhouettes(ndc, faces[:1] if faces.dim() == 3 else faces, self.image_size, False, True,
                                         self.near, self.far)


class PoseOptimizer(nn.Module):
    """Drop-in for /root/reference/homan/pose_optimization.py:37-160. forward() is differentiable through the
    homan_b200 autograd ops, so the reference's own optimisation loop runs on it unchanged; find_optimal_pose
    below uses the fused PoseFitEngine instead."""

    def __init__(self, ref_image, vertices, faces, textures, rotation_init, translation_init,
                 num_initializations=1, kernel_size=7, K=None, power=0.25, lw_chamfer=0):
        assert ref_image.shape[0] == ref_image.shape[1], "Must be square."
        super().__init__()
        dev = torch.device("cuda")
        vertices = torch.as_tensor(vertices).float().to(dev)
        faces = torch.as_tensor(faces).to(dev)
        self.register_buffer("vertices", vertices.view(1, -1, 3).repeat(num_initializations, 1, 1))
        self.register_buffer("faces", faces.view(1, -1, 3).repeat(num_initializations, 1, 1))
        ref = np.asarray(ref_image)
        self.register_buffer("image_ref", torch.from_numpy((ref > 0).astype(np.float32)).to(dev).repeat(num_initializations, 1, 1))
        self.register_buffer("keep_mask", torch.from_numpy((ref >= 0).astype(np.float32)).to(dev).repeat(num_initializations, 1, 1))
        self.ref_image = ref
        F_ = self.faces.shape[1]
        tex = torch.ones(F_, 1, 1, 1, 3) if textures is None else torch.as_tensor(textures).float()
        self.register_buffer("textures", tex.to(dev).view(1, F_, *tex.shape[-4:]).repeat(num_initializations, 1, 1, 1, 1, 1))
        self.rotations = nn.Parameter(torch.as_tensor(rotation_init).clone().float().to(dev), requires_grad=True)
        translation_init = torch.as_tensor(translation_init).float().to(dev)
        if self.rotations.shape[0] != translation_init.shape[0]:
            translation_init = translation_init.repeat(num_initializations, 1, 1)
        self.translations = nn.Parameter(translation_init.clone().float(), requires_grad=True)
        if K is None:
            K = torch.tensor([[[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1]]], dtype=torch.float32, device=dev)
        self.K = torch.as_tensor(K).float().to(dev)
        self.renderer = _Renderer(ref.shape[0], self.K)
        self.lw_chamfer = lw_chamfer
        # edge / distance-transform term (pose_optimization.py:74-88): edges of the reference mask by a max-pool, their
        # Euclidean distance transform to the power 2 * power, once, on the host (scipy, as upstream)
        self.pool = torch.nn.MaxPool2d(kernel_size=kernel_size, stride=1, padding=kernel_size // 2)
        from scipy.ndimage import distance_transform_edt
        mask_edge = self.compute_edges(self.image_ref[:1]).cpu().numpy()
        edt = distance_transform_edt(1 - (mask_edge > 0)) ** (power * 2)
        self.register_buffer("edt_ref_edge", torch.from_numpy(edt).float().to(dev).repeat(num_initializations, 1, 1))

    def apply_transformation(self):
        return torch.matmul(self.vertices, rot6d_to_matrix(self.rotations)) + self.translations

    def compute_edges(self, silhouette):
        """pose_optimization.py:136-137."""
        return self.pool(silhouette) - silhouette

    def render(self):
        """pose_optimization.py:153-160: RGB renders [N,S,S,3] of the candidates at their current poses."""
        from .shims.neural_renderer import Renderer
        r = self.renderer
        full = Renderer(image_size=r.image_size, K=r.K, R=r.R, t=r.t, orig_size=1, anti_aliasing=False)
        with torch.no_grad():
            images = full.render(self.apply_transformation(), self.faces, torch.tanh(self.textures))[0]
        return images.detach().cpu().numpy().transpose(0, 2, 3, 1)

    def compute_offscreen_loss(self, verts):
        r = self.renderer
        proj = ops.project(verts, r.K, r.R, r.t, r.dist_coeffs, 1)
        coord_xy, coord_z = proj[:, :, :2], proj[:, :, 2:]
        zeros = torch.zeros_like(coord_z)
        lower_right = torch.max(coord_xy - 1, zeros).sum(dim=(1, 2))
        upper_left = torch.max(-1 - coord_xy, zeros).sum(dim=(1, 2))
        behind = torch.max(-coord_z, zeros).sum(dim=(1, 2))
        too_far = torch.max(coord_z - r.far, zeros).sum(dim=(1, 2))
        return lower_right + upper_left + behind + too_far

    def forward(self):
        verts = self.apply_transformation()
        image = self.keep_mask * self.renderer(verts, self.faces, mode="silhouettes")
        loss_dict = {"mask": torch.sum((image - self.image_ref) ** 2, dim=(1, 2))}
        with torch.no_grad():
            a, b = image.detach(), self.image_ref
            iou = (a * b).sum((1, 2)) / ((a + b).clamp(0, 1).sum((1, 2)) + 1e-6)
        if self.lw_chamfer != 0:   # (0 in every call of the reference itself; the pool and the product are torch ops)
            loss_dict["chamfer"] = self.lw_chamfer * torch.sum(self.compute_edges(image) * self.edt_ref_edge, dim=(1, 2))
        else:
            loss_dict["chamfer"] = torch.zeros_like(loss_dict["mask"])
        loss_dict["offscreen"] = OFFSCREEN_WEIGHT * self.compute_offscreen_loss(verts)
        return loss_dict, iou, image


def find_optimal_pose(vertices, faces, mask, bbox, square_bbox, image_size, K=None, num_iterations=50,
                      num_initializations=2000, lr=1e-2, image=None, debug=True, viz_folder="tmp", viz_step=10,
                      sort_best=True, rotations_init=None, viz=True, return_engine=False):
    """/root/reference/homan/pose_optimization.py:219-383 (debug plots and the progress bar are not reproduced).
    Returns a PoseOptimizer whose rotations / translations are the fitted candidates."""
    dev = torch.device("cuda")
    vertices = torch.as_tensor(vertices).float().to(dev)
    faces = torch.as_tensor(faces).to(dev)
    x, y, b, _ = square_bbox
    K_t = torch.as_tensor(np.asarray(K) if not torch.is_tensor(K) else K).float()
    camintr_roi = get_K_crop_resize(K_t.view(1, 3, 3).cpu(), torch.tensor([[x, y, x + b, y + b]], dtype=torch.float32),
                                    [REND_SIZE]).to(dev)
    K_dev = K_t.view(1, 3, 3).to(dev)
    if rotations_init is None:
        rotations_init = compute_random_rotations(num_initializations, upright=False, device=dev)
    rotations_init = torch.as_tensor(rotations_init).float().to(dev)
    translations_init = TCO_init_from_boxes_zup_autodepth(bbox, torch.matmul(vertices.unsqueeze(0), rotations_init),
                                                          K_dev).unsqueeze(1)
    camintr_roi[:, :2] = camintr_roi[:, :2] / REND_SIZE  # crop intrinsics in the unit-image convention of the renderer
    N = rotations_init.shape[0]
    eng = PoseFitEngine(vertices, faces, mask, camintr_roi, matrix_to_rot6d(rotations_init), translations_init, lr=lr)
    eng.fit(num_iterations)
    best_rots, best_trans, best_losses = eng.rotations.clone(), eng.translations.clone(), eng.total.clone()
    if sort_best:
        inds = torch.argsort(best_losses)
        best_trans, best_rots = best_trans[inds][:num_initializations], best_rots[inds][:num_initializations]
        if num_iterations > 0 and math.isfinite(float(eng.best[0])):
            best_rots = torch.cat((eng.best[1:7].view(1, 3, 2), best_rots[:-1]), 0)
            best_trans = torch.cat((eng.best[7:10].view(1, 1, 3), best_trans[:-1]), 0)
    model = PoseOptimizer(ref_image=mask, vertices=vertices, faces=faces, textures=None,
                          rotation_init=best_rots, translation_init=best_trans, num_initializations=N, K=camintr_roi)
    if return_engine:
        return model, eng
    return model


def find_optimal_poses(image_size, faces=None, vertices=None, annotations=None, images=None, Ks=None,
                       num_iterations=50, num_initializations=2000, viz_path="tmp.png", debug=False):
    """/root/reference/homan/pose_optimization.py:386-488: per frame, refine num_initializations candidates
    (initialised from the previous frame's rotations), keep the candidate motion with the best mean IoU."""
    dev = torch.device("cuda")
    vertices = torch.as_tensor(np.asarray(vertices)).float().to(dev)
    faces = torch.as_tensor(np.asarray(faces)).to(dev)
    assert faces.dim() == 2 and faces.shape[1] == 3 and vertices.dim() == 2 and vertices.shape[1] == 3
    previous_rotations, all_params, all_ious = None, [], []
    images = images if images is not None else [None] * len(annotations)
    for image, annotation, K in zip(images, annotations, Ks):
        model, eng = find_optimal_pose(vertices=vertices, faces=faces, image=image, mask=annotation["target_crop_mask"],
                                       bbox=annotation["bbox"], square_bbox=annotation["square_bbox"],
                                       image_size=image_size, K=K, num_iterations=num_iterations,
                                       num_initializations=num_initializations, debug=debug, sort_best=False,
                                       rotations_init=previous_rotations, return_engine=True)
        eng.evaluate()  # `_, iou, _ = model()` at the fitted parameters
        rotations = rot6d_to_matrix(eng.rotations).detach()
        all_params.append({
            "rotations": rotations, "translations": eng.translations.detach().clone(),
            "target_masks": torch.as_tensor(np.asarray(annotation["target_crop_mask"])).to(dev),
            "K_roi": model.K.detach(), "masks": torch.as_tensor(annotation["full_mask"]).to(dev),
            "verts": vertices.detach(), "verts_trans": eng.verts.detach().clone()})
        previous_rotations = rotations
        all_ious.append(eng.iou)
    all_ious = torch.stack(all_ious)
    best_idx = torch.argsort(all_ious.mean(0))[-1]
    out = []
    for params, info in zip(all_params, annotations):
        final = {k: params[k][best_idx].unsqueeze(0) for k in ("rotations", "translations", "verts_trans")}
        for k in ("target_masks", "K_roi", "masks", "verts"):
            final[k] = params[k].unsqueeze(0)
        final["full_mask"] = torch.as_tensor(info["full_mask"]).to(dev)
        out.append(final)
    return out
