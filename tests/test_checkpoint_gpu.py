"""GPU: checkpoint / wire-format compatibility with the reference. tests/golden/ref_state_dict.npz
holds the `joint_fit.pt` payload of the UNMODIFIED reference (HOMan.state_dict() minus `mano_model.*`,
/root/reference/fit_vid_dataset.py:365-372) after a 2-iteration fit, and the losses the reference evaluates at that
state. homan_b200.HOMan must expose the same entries (names, shapes), resume from them the way step 2 of the
reference does (`load_state_dict(strict=False)`, /root/reference/homan/jointopt.py:126-127) and reproduce the losses."""
import os

import numpy as np
import pytest
import torch

from golden_utils import load, reference_inputs

pytestmark = pytest.mark.gpu
STATE = os.path.join(os.path.dirname(__file__), "golden", "ref_state_dict.npz")


def _model(mano_assets, lw):
    from homan_b200.homan import HOMan
    z, batch, _, _ = load("ref_small_step2", mano_assets["right"])
    inp = reference_inputs(batch, 0, mano_assets["right"])
    cat = lambda seq, key: torch.cat([p[key] for p in seq])  # noqa: E731
    pp, op = inp["person_parameters"], inp["object_parameters"]
    return HOMan(
        hand_sides=["right"], translations_object=cat(op, "translations"), rotations_object=cat(op, "rotations"),
        verts_object_og=torch.from_numpy(inp["objvertices"]), faces_object=torch.from_numpy(inp["objfaces"]),
        target_masks_object=cat(op, "target_masks"), target_masks_hand=cat(pp, "target_masks"),
        verts_hand_og=cat(pp, "verts"), ref_verts2d_hand=cat(pp, "verts2d"), mano_trans=cat(pp, "mano_trans"),
        mano_rot=cat(pp, "mano_rot"), mano_pca_pose=cat(pp, "mano_pca_pose"), mano_betas=cat(pp, "mano_betas"),
        translations_hand=cat(pp, "translations"), rotations_hand=cat(pp, "rotations"), faces_hand=pp[0]["faces"],
        masks_object=torch.zeros(4, 8, 8, dtype=torch.bool), masks_hand=cat(pp, "masks"), cams_hand=cat(pp, "cams"),
        camintr_rois_object=torch.cat([o["K_roi"][:, 0] for o in op]), camintr_rois_hand=cat(pp, "K_roi"),
        camintr=inp["camintr"], class_name="default", int_scale_init=1, mano_asset=mano_assets["right"],
        loss_weights=lw)


def test_state_dict_schema_matches_the_reference_checkpoint(mano_assets):
    g = np.load(STATE)
    _, _, lw, _ = load("ref_small_step2", mano_assets["right"])
    sd = _model(mano_assets, lw).state_dict()
    ref_keys = [k[3:] for k in g.files if k.startswith("sd_")]
    assert len(ref_keys) == 29
    for k in ref_keys:
        assert k in sd, f"checkpoint entry {k} missing"
        assert tuple(sd[k].shape) == tuple(g["sd_" + k].shape), (k, tuple(sd[k].shape), g["sd_" + k].shape)
    extra = sorted(set(sd) - set(ref_keys))
    assert all(k.startswith("mano_model") for k in extra) or not extra, extra


def test_resume_from_a_reference_checkpoint_reproduces_its_losses(mano_assets):
    g = np.load(STATE)
    _, _, lw, _ = load("ref_small_step2", mano_assets["right"])
    model = _model(mano_assets, lw)
    state = {}
    for k in g.files:
        if k.startswith("sd_"):
            a = g[k]
            state[k[3:]] = torch.from_numpy(a.astype(np.float32) if a.dtype == np.int8 else a)
    missing, unexpected = model.load_state_dict(state, strict=False)   # jointopt.py:126-127
    assert not unexpected and all(k.startswith("mano_model") for k in missing), (missing, unexpected)
    loss_dict, metric_dict = model(lw)
    for k in g.files:
        if k.startswith("eval_"):
            ref = float(g[k])
            got = float(loss_dict[k[5:]])
            assert abs(got - ref) <= 1e-4 * max(abs(ref), 1e-7) + 1e-9, (k, got, ref)
    assert abs(metric_dict["iou_object"] - float(g["metric_iou_object"])) <= 1e-3
    # and the resumed model keeps optimising through the fused engine (parameters alias its buffers)
    before = model.translations_object.detach().clone()
    model.engine.step()
    assert not torch.equal(before, model.translations_object.detach())


def test_step1_to_indep_fit_pickle_to_joint_fit_and_resume(mano_assets, tmp_path):
    """The sample loop of the reference's driver end to end on a synthetic 2-frame clip (fit_vid_dataset.py:283-372):
    object-pose initialisation (find_optimal_poses) -> `indep_fit.pkl` (the five-key dict, pickle.dump / load) ->
    optimize_hand_object on the loaded dicts -> `joint_fit.pt` (state_dict minus mano_model.*, torch.save / load) -> a
    second call resumed from it starts at the saved state (the reference does not save the optimiser either)."""
    import pickle
    from homan_b200 import pose_optimization as po, synth
    from homan_b200.jointopt import optimize_hand_object
    from homan_b200.workload import gpu_render_fn
    asset = mano_assets["right"]
    T = 2
    clip = synth.make_clip(T, "ellipsoid80", seed=5, mano_asset=asset, render_fn=gpu_render_fn())
    batch = synth.make_batch(clip, synth.make_inits(clip, 1, seed=9))
    batch["T"] = T
    inp = reference_inputs(batch, 0, asset)
    # ---- step 1: object pose from the masks (what the detector / PointRend stage hands over per frame)
    K_pix = np.array([[600.0, 0, 320.0], [0, 600.0, 320.0], [0, 0, 1]], dtype=np.float32)
    annotations = []
    for t in range(T):
        uv = synth.project_np(clip["gt"]["verts_obj"][t].astype(np.float64), K_pix.astype(np.float64))
        x, y, b = synth._square_roi(uv)
        lo, hi = uv.min(0), uv.max(0)
        annotations.append({"target_crop_mask": clip["target_masks_object"][t], "full_mask": torch.zeros(8, 8, dtype=torch.bool),
                            "bbox": np.array([lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]], np.float32),
                            "square_bbox": np.array([x, y, b, b], np.float32)})
    torch.manual_seed(0)
    object_parameters = po.find_optimal_poses((640, 640, 3), faces=clip["obj_faces"], vertices=clip["obj_verts_can"],
                                              annotations=annotations, images=None, Ks=[K_pix] * T, num_iterations=30,
                                              num_initializations=256)
    for t, fp in enumerate(object_parameters):   # the crop intrinsics the synthetic clip was rendered with, up to the
        # half-pixel terms of libyana's get_K_crop_resize (principal point: 0.5 / 256 of the normalised crop)
        assert torch.allclose(fp["K_roi"][0, 0].cpu(), torch.from_numpy(clip["K_roi_obj"][t]), atol=3e-3)
    # ---- indep_fit.pkl
    indep = {"person_parameters": inp["person_parameters"], "object_parameters": object_parameters,
             "obj_verts_can": inp["objvertices"], "obj_faces": inp["objfaces"], "super2d_img_path": "super2d.png"}
    path = tmp_path / "indep_fit.pkl"
    with open(path, "wb") as fh:
        pickle.dump(indep, fh)
    with open(path, "rb") as fh:
        loaded = pickle.load(fh)
    assert set(loaded) == set(indep)
    for a, b in zip(loaded["object_parameters"], object_parameters):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    # ---- step 2 on the loaded dicts, then joint_fit.pt as the driver writes it
    _, _, lw, _ = load("ref_small_step2", asset)
    kw = dict(loss_weights=lw, lr=1e-2, viz_folder=str(tmp_path), optimize_mano=True, optimize_mano_beta=True,
              image_size=640, mano_asset=asset, person_parameters=loaded["person_parameters"],
              object_parameters=loaded["object_parameters"], objvertices=loaded["obj_verts_can"],
              objfaces=loaded["obj_faces"], camintr=inp["camintr"])
    model, ev, _ = optimize_hand_object(num_iterations=6, **kw)
    assert len(ev["loss"]) == 6 and np.isfinite(ev["loss"]).all()
    assert ev["loss_sil_obj"][0] < 0.05   # the object starts on its mask: step 1 did its job
    ckpt = tmp_path / "joint_fit.pt"
    torch.save({"state_dict": {k: v.contiguous().cpu() for k, v in model.state_dict().items() if "mano_model" not in k}}, ckpt)
    with torch.no_grad():
        at_save, _ = model(lw)
    state = {k: v.cuda() for k, v in torch.load(ckpt)["state_dict"].items()}
    resumed, ev2, _ = optimize_hand_object(num_iterations=1, state_dict=state, **kw)
    total = sum(float(v) * lw[k.replace("loss", "lw")] for k, v in at_save.items())
    assert abs(ev2["loss"][0] - total) <= 1e-5 * abs(total), (ev2["loss"][0], total)
