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
