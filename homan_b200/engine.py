"""Fused fitting engine: the reference's per-iteration hot path as one fixed sequence of sm_100a kernels.

One `FitEngine.step()` is one iteration of the loop of
/root/reference/homan/jointopt.py:158-192 (zero_grad -> HOMan.forward -> weighted sum -> backward ->
Adam) for P independent problems (clips x random inits) of T frames each, i.e. B = P*T images; every
normaliser of the reference losses is evaluated per problem, so problem p follows exactly the
trajectory the reference produces when given problem p alone.

All state lives in caller-visible torch CUDA tensors; kernels are reached through the C ABI
(include/homan_b200.h) on torch's current stream, and the whole iteration is captured in a CUDA graph.
Scope: one hand per frame, hand_proj_mode="persp", optimize_mano=True, optimize_mano_beta=True,
optimize_object_scale=False (the README configuration of the reference, README.md:207-238).
"""
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import call, current_stream, ptr

NPART = 16
PART = {"smooth_hand": 0, "smooth_obj": 1, "v2d_hand": 2, "v2d_px": 3, "inter": 4, "pca": 5, "sil_obj": 6,
        "iou_obj": 7, "sil_hand": 8, "iou_hand": 9, "collision": 10, "contact": 11, "mindist": 12, "inter_flag": 13}
VL_SMOOTH, VL_V2D, VL_INTER, VL_PCA = 1, 2, 4, 8
LOSS_SLOTS = {"loss_pca": "pca", "loss_smooth_obj": "smooth_obj", "loss_smooth_hand": "smooth_hand",
              "loss_collision": "collision", "loss_contact": "contact", "loss_v2d_hand": "v2d_hand",
              "loss_sil_obj": "sil_obj", "loss_sil_hand": "sil_hand", "loss_inter": "inter"}
METRIC_SLOTS = {"v2d_hand": "v2d_px", "iou_object": "iou_obj", "iou_hand": "iou_hand", "handobj_maxdist": "mindist"}
PARAM_ORDER = ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
               "mano_pca_pose", "mano_betas", "mano_rot", "mano_trans")
REND_SIZE = 256            # /root/reference/homan/constants.py:32
SDF_GRID = 32              # /root/reference/homan/interactions/scenesdf.py:14
SDF_SCALE_FACTOR = 0.2     # /root/reference/homan/interactions/scenesdf.py:77
COLLISION_THRESH = 0.020   # /root/reference/homan/interactions/contactloss.py:156
MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)


def mano_blob(asset, ncomps=16, device="cuda"):
    """Packs a MANO asset (dict with the MANO_RIGHT.pkl keys) into the fp32 blob of hm_mano_fwd."""
    a = {k: np.asarray(v) for k, v in asset.items()}
    parents = a["parents"] if "parents" in a else np.asarray(a["kintree_table"])[0]
    parents = [int(x) for x in parents]
    if tuple(parents[1:]) != MANO_PARENTS[1:]:
        raise _lib.HomanB200Error("unexpected MANO kinematic tree")
    vt = a["v_template"].astype(np.float64).reshape(778, 3)
    sd = a["shapedirs"].astype(np.float64).reshape(778, 3, 10)
    pd = a["posedirs"].astype(np.float64)
    if pd.shape == (778, 3, 135):  # original MANO pickle layout
        pd = pd.reshape(778 * 3, 135).T
    pd = pd.reshape(135, 778 * 3)
    jr = a["J_regressor"]
    jr = np.asarray(jr.todense() if hasattr(jr, "todense") else jr, dtype=np.float64).reshape(16, 778)
    w = a["weights"].astype(np.float64).reshape(778, 16)
    comps = a["hands_components"].astype(np.float64)[:ncomps]
    mean = np.zeros(48)
    mean[:45] = a["hands_mean"].astype(np.float64)
    parts = [vt.ravel(), sd.ravel(), pd.ravel(), (jr @ vt).ravel(), np.einsum("jv,vcl->jcl", jr, sd).ravel(),
             w.ravel(), mean, comps.ravel()]
    return torch.from_numpy(np.concatenate(parts).astype(np.float32)).to(device)


class FitEngine:
    def __init__(self, batch, loss_weights, lr=1e-2, mano_asset=None, device="cuda", use_graph=True, ncomps=16,
                 betas=(0.9, 0.999), eps=1e-8, deterministic=False):
        if not torch.cuda.is_available():
            raise _lib.HomanB200Error("FitEngine needs a CUDA device (there is no CPU path)")
        _lib.lib()
        self.device = dev = torch.device(device)
        self.P, self.T = int(batch["obj_t"].shape[0]), int(batch["obj_t"].shape[1])
        self.B = B = self.P * self.T
        self.lw = {k: float(v) for k, v in loss_weights.items()}
        self.lr, self.adam_betas, self.adam_eps = float(lr), betas, float(eps)
        self.image_size = float(batch.get("image_size", 640))
        self.side_left = 1 if batch.get("side", "right") == "left" else 0
        self.ncomps = ncomps
        asset = mano_asset if mano_asset is not None else batch["mano_asset"]
        self.mano = mano_blob(asset, ncomps, dev)

        # ---- constants (shapes from the batch; values come in through upload())
        # One object mesh for every problem ([Vo,3] / [Fo,3]), or one per clip ([C,Vo,3] / [C,Fo,3] + "clip_of_problem"
        # [P]: the reference fits a different object per sample, fit_vid_dataset.py:190-296). The clips of one batch
        # must agree on Vo and Fo; the per-clip tables are expanded to one mesh / face list per image on the device.
        ov = np.asarray(batch["obj_verts_can"])
        self.n_meshes = 1 if ov.ndim == 2 else int(ov.shape[0])
        self.Vo = Vo = int(ov.shape[-2])
        Fo = int(np.asarray(batch["obj_faces"]).shape[-2])
        R = REND_SIZE
        nm = 1 if self.n_meshes == 1 else B
        self.mesh_obj = torch.empty(nm, Vo, 3, device=dev)
        self.faces_obj = torch.empty(nm, Fo, 3, dtype=torch.int32, device=dev)
        self.faces_hand = torch.as_tensor(np.ascontiguousarray(batch["hand_faces"]).astype(np.int32)).to(dev).view(1, -1, 3)
        self.faces_hand_closed = torch.as_tensor(np.ascontiguousarray(asset["closed_faces"]).astype(np.int32)).to(dev).view(-1, 3)
        self.camintr = torch.empty(B, 3, 3, device=dev)
        self.K_roi_obj = torch.empty(B, 3, 3, device=dev)
        self.K_roi_hand = torch.empty(B, 3, 3, device=dev)
        self.ref_verts2d = torch.empty(B, 778, 2, device=dev)
        self.target_obj = torch.empty(B, R, R, dtype=torch.int8, device=dev)
        self.target_hand = torch.empty(B, R, R, dtype=torch.int8, device=dev)
        self.norm_obj = torch.empty(B, device=dev)
        self.norm_hand = torch.empty(B, device=dev)
        self.scale_obj = torch.ones(1, device=dev)
        self.scale_hand = torch.ones(1, device=dev)

        # ---- parameters: one flat buffer (fused Adam), reference names as views
        D = int(np.asarray(batch["pca"]).shape[-1])
        self.pca_dim = D
        shapes = {"translations_object": (B, 1, 3), "rotations_object": (B, 3, 2), "translations_hand": (B, 1, 3),
                  "rotations_hand": (B, 3, 2), "mano_pca_pose": (B, D), "mano_betas": (B, 10), "mano_rot": (B, 3),
                  "mano_trans": (B, 3)}
        lrs = {"translations_object": lr, "rotations_object": 10 * lr, "translations_hand": lr,
               "rotations_hand": 10 * lr, "mano_pca_pose": 10 * lr, "mano_betas": 10 * lr,
               "mano_rot": 0.0, "mano_trans": 0.0}   # jointopt.py:128-151 (mano_rot / mano_trans match no group)
        n = sum(int(np.prod(s)) for s in shapes.values())
        self.n_params = n
        self.flat = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.lr_elem = torch.zeros(n, device=dev)
        self.params, self.grads, off = {}, {}, 0
        self._segments = {}
        for k in PARAM_ORDER:
            m = int(np.prod(shapes[k]))
            self._segments[k] = (off, m, shapes[k])
            self.params[k] = self.flat[off:off + m].view(shapes[k])
            self.lr_elem[off:off + m] = lrs[k]
            off += m

        # ---- per-iteration scratch; everything that must start at zero lives in one region
        Vo = self.Vo
        sizes = {"grads": n, "g_verts_obj": B * Vo * 3, "g_verts_hand": B * 778 * 3, "g_ndc_obj": B * Vo * 3,
                 "g_ndc_hand": B * 778 * 3, "partials": B * NPART, "g_cdet": B * 3}
        self.zero_region = torch.zeros(sum(sizes.values()), device=dev)
        z, off = {}, 0
        for k, m in sizes.items():
            z[k] = self.zero_region[off:off + m]
            off += m
        self.grad_flat = z["grads"]
        for k in PARAM_ORDER:
            o, m, shp = self._segments[k]
            self.grads[k] = self.grad_flat[o:o + m].view(shp)
        self.g_verts_obj = z["g_verts_obj"].view(B, Vo, 3)
        self.g_verts_hand = z["g_verts_hand"].view(B, 778, 3)
        self.g_ndc_obj = z["g_ndc_obj"].view(B, Vo, 3)
        self.g_ndc_hand = z["g_ndc_hand"].view(B, 778, 3)
        self.partials = z["partials"].view(B, NPART)
        self.g_cdet = z["g_cdet"].view(B, 3)
        self.verts_obj = torch.empty(B, Vo, 3, device=dev)
        self.verts_hand = torch.empty(B, 778, 3, device=dev)
        self.vposed = torch.empty(B, 778 * 3, device=dev)   # posed template of the iteration (hm_mano_fwd -> hm_mano_bwd)
        self.ndc_obj = torch.empty(B, Vo, 3, device=dev)
        self.ndc_hand = torch.empty(B, 778, 3, device=dev)
        self.losses = torch.zeros(self.P, NPART, device=dev)
        self.total = torch.zeros(self.P, device=dev)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=dev)

        # test mode: the scattered gradient sums (raster backward, contact) go through 64-bit fixed-point accumulators,
        # which makes an iteration bit-reproducible whatever the arrival order of CTAs and warps
        self.deterministic = bool(deterministic)
        self.overlap_streams, self._side_streams = True, None
        self.fused_prep = True   # hm_sil_loss_prep instead of hm_sil_loss_fwd_bwd + hm_raster_grad_prep (same outputs)
        self.fixed_region = None
        if self.deterministic:
            self.fixed_region = torch.zeros(2 * B * Vo * 3 + B * 778 * 3, dtype=torch.int64, device=dev)
            o = B * Vo * 3
            self.fx_ndc_obj, self.fx_vobj = self.fixed_region[:o], self.fixed_region[o:2 * o]
            self.fx_ndc_hand = self.fixed_region[2 * o:2 * o + B * 778 * 3]
        self.configure(loss_weights)
        self.gpu_launches_per_step = 0
        self.graph = None
        self.use_graph = use_graph
        self.iteration = 0
        self.upload(self.stage_host(batch, pin=False))


    def configure(self, loss_weights):
        """(Re)selects the active loss terms (gating of /root/reference/homan/homan.py:433-506: a term runs iff its
        weight is > 0) without touching parameters or optimiser state. Invalidates the captured graph."""
        self.lw = {k: float(v) for k, v in loss_weights.items()}
        B, Vo, R, dev = self.B, self.Vo, REND_SIZE, self.device
        on = lambda k: self.lw.get(k, 0.0) > 0  # noqa: E731  (gating of homan.py:433-506)
        self.on_sil_obj, self.on_sil_hand = on("lw_sil_obj"), on("lw_sil_hand")
        self.on_smooth = on("lw_smooth_hand") or on("lw_smooth_obj")
        self.on_v2d, self.on_inter, self.on_pca = on("lw_v2d_hand"), on("lw_inter"), on("lw_pca")
        self.on_contact, self.on_collision = on("lw_contact"), on("lw_collision")
        if self.on_sil_obj and not hasattr(self, "rb_obj"):
            self.rb_obj = ops.RasterBuffers(B, Vo, self.faces_obj.shape[1], R, True, dev)
            self.ga_obj = torch.empty(B, R, R, device=dev)
        if self.on_sil_hand and not hasattr(self, "rb_hand"):
            self.rb_hand = ops.RasterBuffers(B, 778, self.faces_hand.shape[1], R, True, dev)
            self.ga_hand = torch.empty(B, R, R, device=dev)
        if self.on_collision and not hasattr(self, "phi_scratch"):
            self.phi_scratch = torch.empty(B, SDF_GRID ** 3, device=dev)
        w = torch.zeros(NPART)
        for name, slot in LOSS_SLOTS.items():
            w[PART[slot]] = self.lw.get(name.replace("loss_", "lw_"), 0.0)
        self.weights_part = w.to(dev)
        # constant terms of the total: scale priors are (1 - 1)^2 = 0 while the scales are buffers
        self.const_losses = {}
        if on("lw_scale_obj"):
            self.const_losses["loss_scale_obj"] = 0.0
        if on("lw_scale_hand"):
            self.const_losses["loss_scale_hand"] = 0.0
        self.graph = None

    # ------------------------------------------------------------------ host -> device
    @staticmethod
    def stage_host(batch, pin=True):
        """Host-side staging of one problem batch: contiguous (optionally pinned) torch tensors in the
        layouts upload() copies from. Masks travel as int8 {-1, 0, 1}."""
        def t(x, dtype):
            x = torch.as_tensor(np.ascontiguousarray(x)).to(dtype).contiguous()
            return x.pin_memory() if pin and torch.cuda.is_available() else x
        P, T = np.asarray(batch["obj_t"]).shape[:2]
        f = torch.float32
        mesh, faces = np.asarray(batch["obj_verts_can"]), np.asarray(batch["obj_faces"])
        if mesh.ndim == 3:   # one object per clip -> one per image (problem-major)
            cop = np.asarray(batch["clip_of_problem"]).astype(np.int64)
            if cop.shape != (P,) or faces.ndim != 3 or faces.shape[0] != mesh.shape[0]:
                raise _lib.HomanB200Error("per-clip objects need obj_verts_can [C,Vo,3], obj_faces [C,Fo,3], clip_of_problem [P]")
            mesh, faces = np.repeat(mesh[cop], T, 0), np.repeat(faces[cop], T, 0)
        return {
            "mesh_obj": t(mesh, f), "faces_obj": t(faces, torch.int32),
            "camintr": t(batch["camintr"], f), "K_roi_obj": t(batch["K_roi_obj"], f),
            "K_roi_hand": t(batch["K_roi_hand"], f), "ref_verts2d": t(batch["verts2d"], f),
            "target_obj": t(batch["target_masks_object"], torch.int8),
            "target_hand": t(batch["target_masks_hand"], torch.int8),
            "translations_object": t(batch["obj_t"], f), "rotations_object": t(np.asarray(batch["obj_R"])[..., :2], f),
            "translations_hand": t(batch["hand_t"], f), "rotations_hand": t(np.asarray(batch["hand_R"])[..., :2], f),
            "mano_pca_pose": t(batch["pca"], f), "mano_rot": t(batch["mano_rot"], f),
            "mano_trans": t(batch["mano_trans"], f),
        }

    def upload(self, host):
        """Copies a staged batch into the engine's device buffers (async on the current stream), resets the
        optimiser state. Returns the number of bytes copied host -> device."""
        nbytes = 0
        dst = {"mesh_obj": self.mesh_obj, "faces_obj": self.faces_obj, "camintr": self.camintr,
               "K_roi_obj": self.K_roi_obj, "K_roi_hand": self.K_roi_hand, "ref_verts2d": self.ref_verts2d,
               "target_obj": self.target_obj, "target_hand": self.target_hand}
        dst.update({k: self.params[k] for k in PARAM_ORDER if k != "mano_betas"})
        for k, d in dst.items():
            src = host[k]
            if src.numel() != d.numel():
                raise _lib.HomanB200Error(f"upload: {k} has {src.numel()} elements, engine expects {d.numel()}")
            d.copy_(src.view(d.shape), non_blocking=True)
            nbytes += src.numel() * src.element_size()
        self.params["mano_betas"].zero_()   # re-zeroed by the reference: homan/homan.py:108
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_counter.zero_()
        R = REND_SIZE
        keep_o = (self.target_obj >= 0).view(self.P, -1).sum(1).float()       # per problem (losses.py:189-190)
        self.norm_obj.copy_((1.0 / (keep_o * self.T)).repeat_interleave(self.T))
        keep_h = (self.target_hand >= 0).view(self.B, R * R).sum(1).float()   # per image (intended sil_hand)
        self.norm_hand.copy_(1.0 / (keep_h * self.T))
        self.iteration = 0
        return nbytes

    # ------------------------------------------------------------------ one iteration, kernel by kernel
    def _silhouette(self, verts, K_roi, faces, rb, target, norm, weight, ga, g_ndc, g_verts, slot_loss, slot_iou, s,
                    fixed=None):
        B, V = verts.shape[:2]
        ndc = self.ndc_obj if verts is self.verts_obj else self.ndc_hand
        call("hm_project_fwd", ptr(verts), ptr(K_roi), B, None, None, None, 0, 1.0, 1e-9, B, V, ptr(ndc), s)
        ops.raster_forward(rb, ndc, faces)
        pb = self.partials.data_ptr()
        if self.fused_prep:   # loss + gradient + sweep masks / run lists in one kernel
            call("hm_sil_loss_prep", ptr(rb.alpha), ptr(target), ptr(norm), weight, B, REND_SIZE, 1,
                 pb + 4 * PART[slot_loss], NPART, pb + 4 * PART[slot_iou], NPART, ptr(ga), ptr(rb.cov_row),
                 ptr(rb.cov_col), ptr(rb.m_row), ptr(rb.m_col), ptr(rb.runs), ptr(rb.run_counts), s)
        else:
            call("hm_sil_loss_fwd_bwd", ptr(rb.alpha), ptr(target), ptr(norm), weight, B, REND_SIZE,
                 pb + 4 * PART[slot_loss], NPART, pb + 4 * PART[slot_iou], NPART, ptr(ga), s)
        ops.raster_backward(rb, ga, g_ndc, grad_fixed=fixed, prepared=self.fused_prep)
        return (5 if self.fused_prep else 7) + (1 if fixed is not None else 0)

    def _silhouette_finish(self, verts, K_roi, g_ndc, g_verts, s):
        """d loss / d NDC -> d loss / d vertices, accumulated into the vertex gradient the other terms also write."""
        B, V = verts.shape[:2]
        call("hm_project_bwd", ptr(verts), ptr(K_roi), B, None, None, 1.0, 1e-9, B, V, ptr(g_ndc), ptr(g_verts), 1, s)
        return 1

    def _forward_vertices(self, s):
        p = self.params
        call("hm_rigid_fwd", ptr(self.mesh_obj), self.mesh_obj.shape[0], ptr(p["rotations_object"]), ptr(p["translations_object"]),
             ptr(self.scale_obj), self.B, self.Vo, ptr(self.verts_obj), s)
        call("hm_mano_fwd", ptr(self.mano), self.ncomps, self.side_left, ptr(p["mano_pca_pose"]), self.pca_dim,
             ptr(p["mano_rot"]), ptr(p["mano_betas"]), ptr(p["mano_trans"]), ptr(p["rotations_hand"]),
             ptr(p["translations_hand"]), ptr(self.scale_hand), self.B, ptr(self.verts_hand), None, ptr(self.vposed), s)
        return 2

    def _iteration(self, adam=True):
        s = current_stream()
        lw, B, T = self.lw, self.B, self.T
        n = 1
        self.zero_region.zero_()
        if self.deterministic:
            self.fixed_region.zero_()
            n += 1
        main = torch.cuda.current_stream()
        fork = self.overlap_streams and torch.cuda.is_current_stream_capturing()
        if fork and self._side_streams is None:
            # the hand chain ends in the MANO backward, a launch that leaves most of the GPU idle: it gets the higher
            # priority so that it finishes first and that tail runs under the object chain's last kernels
            hp = int(os.environ.get("HOMAN_B200_HAND_PRIORITY", "-1"))
            self._side_streams = [torch.cuda.Stream(), torch.cuda.Stream(priority=hp)]
        if fork and self.on_sil_obj and self.on_sil_hand:
            # each chain starts from its own vertices: object placement on the first side stream, MANO on the second
            sa, sb = self._side_streams
            p = self.params
            sa.wait_stream(main)
            sb.wait_stream(main)
            with torch.cuda.stream(sa):
                call("hm_rigid_fwd", ptr(self.mesh_obj), self.mesh_obj.shape[0], ptr(p["rotations_object"]),
                     ptr(p["translations_object"]), ptr(self.scale_obj), self.B, self.Vo, ptr(self.verts_obj), current_stream())
                ev_obj = torch.cuda.Event()
                ev_obj.record(sa)
            with torch.cuda.stream(sb):
                call("hm_mano_fwd", ptr(self.mano), self.ncomps, self.side_left, ptr(p["mano_pca_pose"]), self.pca_dim,
                     ptr(p["mano_rot"]), ptr(p["mano_betas"]), ptr(p["mano_trans"]), ptr(p["rotations_hand"]),
                     ptr(p["translations_hand"]), ptr(self.scale_hand), self.B, ptr(self.verts_hand), None,
                     ptr(self.vposed), current_stream())
                ev_hand = torch.cuda.Event()
                ev_hand.record(sb)
            main.wait_event(ev_obj)
            main.wait_event(ev_hand)
            n += 2
            verts_on_sides = True
        else:
            n += self._forward_vertices(s)
            verts_on_sides = False
        # The two silhouette chains (projection -> raster forward -> loss -> raster backward) only write their own
        # buffers (NDC gradients, loss slots), so inside a CUDA-graph capture they run on two side streams next to the
        # vertex-space terms on the main stream (fork / join: parallel branches of the graph); the projection backward,
        # which accumulates into the shared vertex gradients, follows the join. Eager launches stay on one stream.
        chains = []
        if self.on_sil_obj:
            chains.append((self.verts_obj, self.K_roi_obj, self.faces_obj, self.rb_obj, self.target_obj, self.norm_obj,
                           lw["lw_sil_obj"], self.ga_obj, self.g_ndc_obj, self.g_verts_obj, "sil_obj", "iou_obj",
                           self.fx_ndc_obj if self.deterministic else None))
        if self.on_sil_hand:
            chains.append((self.verts_hand, self.K_roi_hand, self.faces_hand, self.rb_hand, self.target_hand,
                           self.norm_hand, lw["lw_sil_hand"], self.ga_hand, self.g_ndc_hand, self.g_verts_hand,
                           "sil_hand", "iou_hand", self.fx_ndc_hand if self.deterministic else None))
        if fork:
            for side, c in zip(self._side_streams, chains):
                if not verts_on_sides:
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    n += self._silhouette(*c[:12], current_stream(), c[12])
        else:
            for c in chains:
                n += self._silhouette(*c[:12], s, c[12])
        flags = (VL_SMOOTH if self.on_smooth else 0) | (VL_V2D if self.on_v2d else 0) | \
                (VL_INTER if self.on_inter else 0) | (VL_PCA if self.on_pca else 0)
        if flags:
            call("hm_vertex_losses", ptr(self.verts_hand), ptr(self.verts_obj), ptr(self.camintr),
                 ptr(self.ref_verts2d), ptr(self.params["mano_pca_pose"]), self.pca_dim, B, T, self.Vo,
                 self.image_size, lw.get("lw_smooth_hand", 0.0), lw.get("lw_smooth_obj", 0.0),
                 lw.get("lw_v2d_hand", 0.0), lw.get("lw_inter", 0.0), lw.get("lw_pca", 0.0), flags,
                 ptr(self.partials), ptr(self.g_verts_hand), ptr(self.g_verts_obj), ptr(self.g_cdet),
                 ptr(self.grads["mano_pca_pose"]), s)
            n += 1
        if self.on_contact or self.on_inter:
            call("hm_contact_fwd_bwd", ptr(self.verts_hand), ptr(self.verts_obj), B, T, self.Vo, COLLISION_THRESH,
                 lw.get("lw_contact", 0.0) if self.on_contact else 0.0, ptr(self.partials), ptr(self.g_verts_hand),
                 ptr(self.g_verts_obj), ptr(self.fx_vobj) if self.deterministic else None, s)
            n += 1
            if self.deterministic:
                call("hm_fold_fixed", ptr(self.fx_vobj), self.fx_vobj.numel(), ptr(self.g_verts_obj), s)
                n += 1
        if self.on_collision:
            # pair (hand grid <- object vertices): value only (the object is detached, homan.py:445-449)
            call("hm_sdf_pair", ptr(self.verts_hand), ptr(self.faces_hand_closed), 1, ptr(self.verts_obj), B, 778,
                 self.faces_hand_closed.shape[0], self.Vo, SDF_GRID, SDF_SCALE_FACTOR, 0.0, ptr(self.phi_scratch),
                 ptr(self.partials), None, None, s)
            # pair (object grid <- hand vertices): gradient to the hand
            call("hm_sdf_pair", ptr(self.verts_obj), ptr(self.faces_obj), self.faces_obj.shape[0], ptr(self.verts_hand), B, self.Vo,
                 self.faces_obj.shape[1], 778, SDF_GRID, SDF_SCALE_FACTOR, lw["lw_collision"], ptr(self.phi_scratch),
                 ptr(self.partials), ptr(self.g_verts_hand), None, s)
            n += 2
        g = self.grads
        obj_tail = None   # side stream that carries the object's projection / placement backward, if any
        if fork and os.environ.get("HOMAN_B200_TAIL_OVERLAP", "1") != "0":
            # join per chain: the object's tail (projection backward -> rigid backward) stays on its side stream behind
            # the vertex-space terms of the main stream, the hand's tail (projection backward -> MANO backward) runs on
            # the main stream as soon as the hand chain is done; they meet again before the loss reduction
            ev_terms = torch.cuda.Event()
            ev_terms.record(main)
            for side, c in zip(self._side_streams, chains):
                if c[10] == "sil_obj":
                    side.wait_event(ev_terms)
                    with torch.cuda.stream(side):
                        n += self._silhouette_finish(c[0], c[1], c[8], c[9], current_stream())
                        call("hm_rigid_bwd", ptr(self.mesh_obj), self.mesh_obj.shape[0], ptr(self.params["rotations_object"]),
                             ptr(self.scale_obj), B, self.Vo, ptr(self.g_verts_obj), ptr(g["rotations_object"]),
                             ptr(g["translations_object"]), current_stream())
                    obj_tail = side
                else:
                    main.wait_stream(side)
                    n += self._silhouette_finish(c[0], c[1], c[8], c[9], s)
        else:
            if fork:
                for side, _ in zip(self._side_streams, chains):
                    main.wait_stream(side)
            for c in chains:
                n += self._silhouette_finish(c[0], c[1], c[8], c[9], s)
        call("hm_mano_bwd", ptr(self.mano), self.ncomps, self.side_left, ptr(self.params["mano_pca_pose"]),
             self.pca_dim, ptr(self.params["mano_rot"]), ptr(self.params["mano_betas"]),
             ptr(self.params["mano_trans"]), ptr(self.params["rotations_hand"]), ptr(self.params["translations_hand"]),
             ptr(self.scale_hand), B, ptr(self.vposed), ptr(self.g_verts_hand), ptr(self.g_cdet) if self.on_inter else None,
             ptr(g["mano_pca_pose"]), ptr(g["mano_rot"]), ptr(g["mano_betas"]), ptr(g["mano_trans"]),
             ptr(g["rotations_hand"]), ptr(g["translations_hand"]), s)
        if obj_tail is None:
            call("hm_rigid_bwd", ptr(self.mesh_obj), self.mesh_obj.shape[0], ptr(self.params["rotations_object"]), ptr(self.scale_obj), B,
                 self.Vo, ptr(self.g_verts_obj), ptr(g["rotations_object"]), ptr(g["translations_object"]), s)
        else:
            main.wait_stream(obj_tail)
        call("hm_finalize_losses", ptr(self.partials), ptr(self.weights_part), self.P, T, ptr(self.losses),
             ptr(self.total), ptr(self.step_counter) if adam else None, s)   # forward-only: Adam's step count untouched
        n += 3
        if adam:
            call("hm_adam_step", ptr(self.flat), ptr(self.grad_flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                 ptr(self.lr_elem), self.n_params, self.adam_betas[0], self.adam_betas[1], self.adam_eps,
                 ptr(self.step_counter), s)
            n += 1
        self.gpu_launches_per_step = n
        return n

    def evaluate(self):
        """Forward + backward at the current parameters, no Adam update. Leaves losses / grads filled."""
        self._iteration(adam=False)
        return self.loss_dict()

    def capture(self):
        """Warm up once per kernel (function attributes, lazy module load) on a scratch copy of the state,
        then capture one iteration in a CUDA graph."""
        saved = (self.flat.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_counter.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        mp = int(os.environ.get("HOMAN_B200_MAIN_PRIORITY", "-1"))
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(priority=mp)):
            self._iteration()
        for dst, src in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.step_counter), saved):
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

    # ------------------------------------------------------------------ results
    def loss_dict(self, losses=None):
        """{loss name: [P] numpy} in the reference's naming (unweighted), plus "total" and metrics."""
        ls = (self.losses if losses is None else losses).detach().cpu().numpy()
        out = {}
        for name, slot in LOSS_SLOTS.items():
            if self.lw.get(name.replace("loss_", "lw_"), 0.0) > 0:
                out[name] = ls[:, PART[slot]].copy()
        if self.on_smooth:
            out["loss_smooth_obj"] = ls[:, PART["smooth_obj"]].copy()
            out["loss_smooth_hand"] = ls[:, PART["smooth_hand"]].copy()
        for k, v in self.const_losses.items():
            out[k] = np.full(self.P, v, np.float32)
        return out

    def metric_dict(self, losses=None):
        ls = (self.losses if losses is None else losses).detach().cpu().numpy()
        out = {}
        if self.on_v2d:
            out["v2d_hand"] = ls[:, PART["v2d_px"]].copy()
        if self.on_sil_obj:
            out["iou_object"] = ls[:, PART["iou_obj"]].copy()
        if self.on_sil_hand:
            out["iou_hand"] = ls[:, PART["iou_hand"]].copy()
        if self.on_inter:
            out["handobj_maxdist"] = ls[:, PART["mindist"]].copy()
        return out

    def fit(self, num_iterations, record=True):
        """Runs the loop. Returns {"losses": {name: [iters, P]}, "total": [iters, P], "params": {...}}."""
        hist = torch.zeros(num_iterations, self.P, NPART, device=self.device) if record else None
        tot = torch.zeros(num_iterations, self.P, device=self.device) if record else None
        for it in range(num_iterations):
            self.step()
            if record:
                hist[it].copy_(self.losses)
                tot[it].copy_(self.total)
        torch.cuda.synchronize()
        out = {"params": {k: v.detach().cpu().numpy().copy() for k, v in self.params.items()}}
        if record:
            h = hist.cpu()
            out["losses"] = {}
            for it_name in list(self.loss_dict(h[0]).keys()):
                out["losses"][it_name] = np.stack([self.loss_dict(h[i])[it_name] for i in range(num_iterations)])
            out["metrics"] = {k: np.stack([self.metric_dict(h[i])[k] for i in range(num_iterations)])
                              for k in self.metric_dict(h[0])}
            out["total"] = tot.cpu().numpy()
        return out

    def best_init(self, clips):
        """argmin over the inits of every clip (problems are clip-major): (best_index [C], best_loss [C])."""
        inits = self.P // clips
        bi = torch.empty(clips, dtype=torch.int32, device=self.device)
        bl = torch.empty(clips, device=self.device)
        call("hm_argmin_over_inits", ptr(self.total), clips, inits, ptr(bi), ptr(bl), current_stream())
        return bi, bl
