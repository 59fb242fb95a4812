"""Drop-in for /root/reference/homan/homan.py::HOMan on the fitting hot path.

Same constructor keywords (homan.py:27-60), same parameter / buffer names (checkpoint compatible,
SURVEY.md §5), `forward(loss_weights) -> (loss_dict, metric_dict)` with the reference's loss names, and the
accessors the driver uses (`get_verts_object()`, `get_verts_hand()`, `verts_object_init`, `verts_hand_init`,
`faces_hand`, `faces_object`, `state_dict()`; fit_vid_dataset.py:365-379,489-493). All arithmetic runs in the
fused CUDA engine (engine.py); the nn.Parameters are views of the engine's flat parameter buffer.

Extension: `frames_per_problem=T` batches P = B / T independent problems (clips x inits) in front of the
reference's frame axis; the default (T = B) is the reference's single-clip semantics.
Supported configuration (README.md:207-238 of the reference): one hand per frame, hand_proj_mode="persp",
optimize_mano=True, optimize_mano_beta=True, optimize_object_scale=False. Anything else raises.
"""
import numpy as np
import torch
from torch import nn

from . import _lib
from .engine import LOSS_SLOTS, PARAM_ORDER, FitEngine
from .shims.mano_layer import load_asset


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _FusedLosses(torch.autograd.Function):
    """Bridges the fused forward+backward to autograd: outputs the unweighted losses (summed over problems);
    backward hands out the parameter gradients the engine already accumulated for sum_k lw_k * loss_k, and
    therefore requires the upstream gradient of loss k to be lw_k (what jointopt.py:180-183,191 produces)."""

    @staticmethod
    def forward(ctx, model, weights, *params):
        eng = model.engine
        eng._iteration(adam=False)   # forward-only: the Adam step counter is not touched
        ctx.model, ctx.weights = model, weights
        names = model._loss_names
        vals = torch.stack([eng.losses[:, model._slot_of[n]].sum() for n in names]) if names else eng.losses.new_zeros(0)
        return vals

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.model
        expect = torch.tensor([ctx.weights[n] for n in model._loss_names], device=grad_out.device)
        if grad_out.numel() and not torch.allclose(grad_out, expect, rtol=1e-5, atol=0):
            raise _lib.HomanB200Error(
                "homan_b200.HOMan: losses must be combined as sum(loss_k * loss_weights['lw_k']) with the "
                "loss_weights passed to forward() (the fused kernels already applied them)")
        return (None, None) + tuple(model.engine.grads[k].clone() for k in PARAM_ORDER)


class HOMan(nn.Module):
    def __init__(self, translations_object, rotations_object, verts_object_og, faces_object, translations_hand,
                 rotations_hand, verts_hand_og, ref_verts2d_hand, hand_sides, mano_trans, mano_rot, mano_betas,
                 mano_pca_pose, faces_hand, masks_object, masks_hand, camintr_rois_object, camintr_rois_hand,
                 target_masks_object, target_masks_hand, class_name, cams_hand=None, int_scale_init=1.0,
                 camintr=None, optimize_object_scale=False, optimize_ortho_cam=True, hand_proj_mode="persp",
                 optimize_mano=True, optimize_mano_beta=True, inter_type="centroid", image_size=640,
                 frames_per_problem=None, mano_asset=None, mano_root="extra_data/mano", loss_weights=None, lr=1e-2):
        super().__init__()
        if len(hand_sides) != 1:
            raise NotImplementedError("homan_b200.HOMan: one hand per frame")
        if hand_proj_mode != "persp" or not optimize_mano or not optimize_mano_beta or optimize_object_scale:
            raise NotImplementedError("homan_b200.HOMan: persp projection, optimize_mano, optimize_mano_beta, "
                                      "fixed object scale (the reference's README configuration)")
        if inter_type != "centroid" or int_scale_init != 1:
            raise NotImplementedError("homan_b200.HOMan: inter_type='centroid', int_scale_init=1")
        self.hand_sides, self.hand_nb, self.hand_proj_mode = hand_sides, 1, hand_proj_mode
        self.optimize_mano, self.optimize_object_scale, self.image_size = optimize_mano, optimize_object_scale, image_size
        self.class_name, self.lr = class_name, lr
        B = int(translations_object.shape[0])
        T = B if frames_per_problem is None else int(frames_per_problem)
        if B % T:
            raise ValueError("batch size must be a multiple of frames_per_problem")
        P = B // T
        side = hand_sides[0]
        asset = mano_asset if mano_asset is not None else load_asset(mano_root, is_right=(side == "right"))
        if "closed_faces" not in asset:
            closed = np.load("local_data/closed_fmano.npy")  # /root/reference/homan/lossutils.py:15
            asset = dict(asset, closed_faces=closed if side == "right" else closed[:, ::-1].copy())
        rot_o, rot_h = _np(rotations_object), _np(rotations_hand)
        verts_og = _np(verts_object_og)
        faces_o = _np(faces_object)
        cam = np.array([[[1, 0, 0.5], [0, 1, 0.5], [0, 0, 1]]], np.float32) if camintr is None else _np(camintr)
        if cam.ndim == 2:
            cam = cam[None]
        cam = np.broadcast_to(cam, (T,) + cam.shape[1:]) if cam.shape[0] == 1 else cam
        rs = lambda x, *tail: np.ascontiguousarray(_np(x), dtype=np.float32).reshape((P, T) + tail)  # noqa: E731
        batch = {
            "side": side, "image_size": image_size, "mano_asset": asset,
            "obj_verts_can": np.ascontiguousarray(verts_og[0] if verts_og.ndim == 3 else verts_og, np.float32),
            "obj_faces": np.ascontiguousarray(faces_o[0] if faces_o.ndim == 3 else faces_o),
            "hand_faces": np.ascontiguousarray(_np(faces_hand).reshape(-1, 3)[:1538]),
            "obj_t": rs(translations_object, 3), "obj_R": rs(rot_o, 3, rot_o.shape[-1]),
            "hand_t": rs(translations_hand, 3), "hand_R": rs(rot_h, 3, rot_h.shape[-1]),
            "pca": rs(mano_pca_pose, _np(mano_pca_pose).shape[-1]), "mano_rot": rs(mano_rot, 3),
            "mano_trans": rs(mano_trans, 3), "betas": rs(mano_betas, 10),
            "camintr": np.ascontiguousarray(np.tile(cam, (P, 1, 1)) if cam.shape[0] == T else cam, np.float32).reshape(P, T, 3, 3),
            "K_roi_obj": rs(camintr_rois_object, 3, 3), "K_roi_hand": rs(camintr_rois_hand, 3, 3),
            "verts2d": rs(ref_verts2d_hand, 778, 2),
            "target_masks_object": rs(target_masks_object, 256, 256), "target_masks_hand": rs(target_masks_hand, 256, 256),
        }
        self._batch, self._asset = batch, asset
        self.P, self.T = P, T
        self._build_engine(loss_weights or {})
        dev = self.engine.device
        buf = lambda x: torch.as_tensor(_np(x)).to(dev)  # noqa: E731
        self.register_buffer("verts_object_og", buf(verts_object_og).float())
        self.register_buffer("verts_hand_og", buf(verts_hand_og).float())
        self.register_buffer("ref_verts2d_hand", buf(ref_verts2d_hand).float())
        self.register_buffer("int_scales_hand", torch.ones(1, device=dev))
        self.register_buffer("int_scales_object", torch.ones(1, device=dev))
        self.register_buffer("int_scale_object_mean", torch.ones(1, device=dev))
        self.register_buffer("int_scale_hand_mean", torch.ones(1, device=dev))
        tmo, tmh = buf(target_masks_object).float(), buf(target_masks_hand).float()
        self.register_buffer("ref_mask_object", (tmo > 0).float())
        self.register_buffer("keep_mask_object", (tmo >= 0).float())
        self.register_buffer("ref_mask_hand", (tmh > 0).float())
        self.register_buffer("keep_mask_hand", (tmh >= 0).float())
        self.register_buffer("camintr_rois_object", buf(camintr_rois_object).float())
        self.register_buffer("camintr_rois_hand", buf(camintr_rois_hand).float())
        self.register_buffer("faces_object", buf(faces_object))
        self.register_buffer("faces_hand", buf(faces_hand))
        self.register_buffer("textures_object", torch.ones(faces_o.shape[0], faces_o.shape[-2], 1, 1, 1, 3, device=dev))
        fh = _np(faces_hand)
        self.register_buffer("textures_hand", torch.ones(fh.shape[0], fh.shape[-2], 1, 1, 1, 3, device=dev))
        self.register_buffer("camintr", buf(cam).float())
        if cams_hand is not None:
            self.cams_hand = nn.Parameter(buf(cams_hand).float(), requires_grad=True)
        if masks_hand is not None:
            self.register_buffer("masks_human", buf(masks_hand))
        mo = buf(masks_object)
        self.register_buffer("masks_object", mo[None] if mo.dim() == 2 else mo)
        with torch.no_grad():
            self.verts_object_init, _ = self.get_verts_object()
            self.verts_hand_init, _ = self.get_verts_hand()
        # ---- visualisation (homan/homan.py:168-217): full-frame renderer, object gold + hand grey, one combined mesh
        from .shims import neural_renderer as nr
        from .visualize import COLORS
        self.renderer = nr.Renderer(image_size=image_size, K=self.camintr.clone(), R=torch.eye(3, device=dev)[None],
                                    t=torch.zeros(1, 3, device=dev), orig_size=1)
        self.renderer.light_direction = [1, 0.5, 1]
        self.renderer.light_intensity_direction = 0.3   # (sic: the reference sets this misspelt, unused attribute)
        self.renderer.light_intensity_ambient = 0.5
        self.renderer.background_color = [1.0, 1.0, 1.0]
        fo_, fh_ = torch.as_tensor(batch["obj_faces"]).long().to(dev), torch.as_tensor(batch["hand_faces"]).long().to(dev)
        self.faces = torch.cat((fo_, fh_ + batch["obj_verts_can"].shape[0]))[None].repeat(B, 1, 1)
        tex = torch.cat((torch.tensor(COLORS["gold"], device=dev).expand(fo_.shape[0], 3),
                         torch.tensor(COLORS["grey"], device=dev).expand(fh_.shape[0], 3)))
        self.textures = tex.view(1, -1, 1, 1, 1, 3).repeat(B, 1, 1, 1, 1, 1)
        # ground-truth overlays (homan/homan.py:200-217): the same topology in green / blue, and prediction + ground truth
        # as one mesh of four parts
        tex_gt = torch.cat((torch.tensor(COLORS["green"], device=dev).expand(fo_.shape[0], 3),
                            torch.tensor(COLORS["blue"], device=dev).expand(fh_.shape[0], 3)))
        self.textures_gt = tex_gt.view(1, -1, 1, 1, 1, 3).repeat(B, 1, 1, 1, 1, 1)
        self.faces_gt = self.faces
        nv = batch["obj_verts_can"].shape[0] + 778
        self.faces_with_gt = torch.cat((self.faces, self.faces + nv), 1)
        self.textures_with_gt = torch.cat((self.textures, self.textures_gt), 1)

    # ------------------------------------------------------------------ engine plumbing
    def _build_engine(self, loss_weights):
        self.engine = FitEngine(self._batch, loss_weights, lr=self.lr, mano_asset=self._asset, use_graph=True)
        for k in PARAM_ORDER:
            # nn.Parameters that alias the engine's flat buffer (fused Adam updates them in place)
            setattr(self, k, nn.Parameter(self.engine.params[k], requires_grad=True))
        self._weights = {k: float(v) for k, v in loss_weights.items()}

    def _sync_weights(self, loss_weights):
        lw = {k: float(v) for k, v in loss_weights.items()}
        if lw != self._weights:
            self.engine.configure(lw)   # parameters (and any optimiser holding them) stay in place
            self._weights = lw

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict=strict)
        for k in PARAM_ORDER:  # keep aliasing the engine buffer after a load
            self.engine.params[k].copy_(getattr(self, k).data.reshape(self.engine.params[k].shape))
            getattr(self, k).data = self.engine.params[k]
        return out

    # ------------------------------------------------------------------ reference API
    def get_verts_object(self, **kwargs):
        s = _lib.current_stream()
        self.engine._forward_vertices(s)
        v = self.engine.verts_obj.clone()
        return v, v

    def get_verts_hand(self, detach_scale=False, **kwargs):
        s = _lib.current_stream()
        self.engine._forward_vertices(s)
        v = self.engine.verts_hand.clone()
        return v, v

    def render_limem(self, renderer, verts, faces, textures, K, max_in_batch=5):
        """homan/homan.py:510-545: (images [N,S,S,3] in [0,1], masks [N,S,S] bool). One batch (memory is not the limit
        here: `max_in_batch` is accepted and ignored)."""
        rgb, _, alpha = renderer.render(vertices=verts, faces=faces, textures=textures, K=K)
        return np.clip(rgb.permute(0, 2, 3, 1).cpu().numpy(), 0, 1), alpha.cpu().numpy().astype(bool)

    def render(self, renderer=None, rotate=False, viz_len=10, max_in_batch=None):
        """homan/homan.py:547-562: RGB render of the fitted object + hand of the first `viz_len` frames (optionally
        rotated about the scene centroid for the "top-down" view)."""
        from .visualize import rot_points
        renderer = self.renderer if renderer is None else renderer
        with torch.no_grad():
            verts = torch.cat((self.get_verts_object()[0], self.get_verts_hand()[0]), 1)
            if rotate:
                verts = rot_points(verts)
            K = renderer.K.view(-1, 3, 3)
            K = K.repeat(verts.shape[0] // K.shape[0], 1, 1) if verts.shape[0] % K.shape[0] == 0 else K[:1].expand(verts.shape[0], -1, -1)
            return self.render_limem(renderer, verts[:viz_len].contiguous(), self.faces[:viz_len], self.textures[:viz_len],
                                     K=K[:viz_len].contiguous(), max_in_batch=max_in_batch)

    def _render_meshes(self, renderer, verts, faces, textures, rotate, viz_len, max_in_batch):
        from .visualize import rot_points
        renderer = self.renderer if renderer is None else renderer
        with torch.no_grad():
            verts = verts.float()
            if rotate:
                verts = rot_points(verts)
            K = renderer.K.view(-1, 3, 3)
            K = K.repeat(verts.shape[0] // K.shape[0], 1, 1) if verts.shape[0] % K.shape[0] == 0 else K[:1].expand(verts.shape[0], -1, -1)
            return self.render_limem(renderer, verts[:viz_len].contiguous(), faces[:viz_len], textures[:viz_len],
                                     K=K[:viz_len].contiguous(), max_in_batch=max_in_batch)

    def _gt_hand(self, verts_hand_gt):
        """The reference passes one [T,778,3] tensor per hand (a list / stacked [H,T,778,3]) or, for one hand, the tensor."""
        v = verts_hand_gt
        if isinstance(v, (list, tuple)):
            if len(v) != 1:
                raise NotImplementedError("homan_b200.HOMan: one hand per frame")
            v = v[0]
        v = torch.as_tensor(_np(v) if not torch.is_tensor(v) else v).to(self.engine.device).float()
        if v.dim() == 4:
            if v.shape[0] != 1:
                raise NotImplementedError("homan_b200.HOMan: one hand per frame")
            v = v[0]
        return v

    def render_gt(self, renderer=None, verts_hand_gt=None, verts_object_gt=None, rotate=False, viz_len=10,
                  max_in_batch=None):
        """homan/homan.py:564-581: the ground-truth object (green) and hand (blue) alone."""
        vo = torch.as_tensor(_np(verts_object_gt) if not torch.is_tensor(verts_object_gt) else verts_object_gt)
        verts = torch.cat((vo.to(self.engine.device).float(), self._gt_hand(verts_hand_gt)), 1)
        return self._render_meshes(renderer, verts, self.faces_gt, self.textures_gt, rotate, viz_len, max_in_batch)

    def render_with_gt(self, renderer=None, verts_hand_gt=None, verts_object_gt=None, rotate=False, viz_len=10, init=False,
                       max_in_batch=None):
        """homan/homan.py:583-613: the fit (or the initialisation, `init`) in gold / grey next to the ground truth in
        green / blue, one render."""
        with torch.no_grad():
            if init:
                vo_p, vh_p = self.verts_object_init, self.verts_hand_init
            else:
                vo_p, vh_p = self.get_verts_object()[0], self.get_verts_hand()[0]
        vo = torch.as_tensor(_np(verts_object_gt) if not torch.is_tensor(verts_object_gt) else verts_object_gt)
        n = min(vo.shape[0], vo_p.shape[0])   # (a model with several problems draws its first clip)
        verts = torch.cat((vo_p[:n], vh_p[:n], vo.to(self.engine.device).float()[:n], self._gt_hand(verts_hand_gt)[:n]), 1)
        return self._render_meshes(renderer, verts, self.faces_with_gt, self.textures_with_gt, rotate, viz_len, max_in_batch)

    def get_joints_hand(self):
        """homan/homan.py:309-339: the 21 hand joints (16 MANO joints + 5 finger tips, reordered) placed in the camera
        frame like the vertices -> (joints [B,21,3], the same) - evaluation helper, not on the fitting path."""
        from .pose_optimization import rot6d_to_matrix
        eng, p = self.engine, self.engine.params
        B = eng.B
        with torch.no_grad():
            verts = torch.empty(B, 778, 3, device=eng.device)
            joints = torch.empty(B, 16, 3, device=eng.device)
            _lib.call("hm_mano_fwd", _lib.ptr(eng.mano), eng.ncomps, eng.side_left, _lib.ptr(p["mano_pca_pose"]), eng.pca_dim,
                      _lib.ptr(p["mano_rot"]), _lib.ptr(p["mano_betas"]), _lib.ptr(p["mano_trans"]), None, None, None, B,
                      _lib.ptr(verts), _lib.ptr(joints), None, _lib.current_stream())
            full = torch.cat((joints, verts[:, [745, 317, 444, 556, 673]]), 1)
            full = full[:, [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]]
            R = rot6d_to_matrix(p["rotations_hand"].reshape(B, 3, 2))
            out = torch.matmul(eng.scale_hand.view(-1, 1, 1) * full, R) + p["translations_hand"].reshape(B, 1, 3)
        return out, out.clone()

    def save_obj(self, fname):
        """homan/homan.py:615-626: the first frame's object + hand as a Wavefront .obj."""
        with torch.no_grad():
            verts = torch.cat((self.get_verts_object()[0][:1], self.get_verts_hand()[0][:1]), 1)[0].cpu().numpy()
        faces = self.faces[0].cpu().numpy()
        with open(fname, "w") as fp:
            for v in verts:
                fp.write(f"v {v[0]:f} {v[1]:f} {v[2]:f}\n")
            for f in faces:
                fp.write(f"f {f[0] + 1:d} {f[1] + 1:d} {f[2] + 1:d}\n")

    def forward(self, loss_weights=None):
        if loss_weights is None:
            loss_weights = self._weights
        self._sync_weights(loss_weights)
        eng = self.engine
        names = [n for n in LOSS_SLOTS if eng.lw.get(n.replace("loss_", "lw_"), 0.0) > 0]
        if eng.on_smooth:
            names += [n for n in ("loss_smooth_obj", "loss_smooth_hand") if n not in names]
        from .engine import PART
        self._loss_names = names
        self._slot_of = {n: PART[LOSS_SLOTS[n]] for n in names}
        weights = {n: eng.lw.get(n.replace("loss_", "lw_"), 0.0) for n in names}
        vals = _FusedLosses.apply(self, weights, *[getattr(self, k) for k in PARAM_ORDER])
        loss_dict = {n: vals[i] for i, n in enumerate(names)}
        for k, v in eng.const_losses.items():
            loss_dict[k] = vals.new_tensor(v)
        metric_dict = {}
        m = eng.metric_dict()
        for k, v in m.items():
            metric_dict[k] = float(v.max() if k == "handobj_maxdist" else v.mean())
        return loss_dict, metric_dict

    def per_problem_losses(self):
        """{loss name: [P]} of the last forward / step (the problem-axis extension)."""
        return self.engine.loss_dict()
