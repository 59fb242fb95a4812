"""CPU: host-side logic of the product (no kernels run here): workload generator, MANO blob packing,
header parsing, and the "no CPU fallback" contract."""
import numpy as np
import pytest
import torch

from homan_b200 import _lib, synth
from homan_b200.engine import FitEngine, mano_blob


def test_synth_is_seeded_and_has_the_reference_schema(mano_assets):
    a = synth.make_clip(3, "ellipsoid80", seed=7, mano_asset=mano_assets["right"])
    b = synth.make_clip(3, "ellipsoid80", seed=7, mano_asset=mano_assets["right"])
    assert np.array_equal(a["gt"]["verts_hand"], b["gt"]["verts_hand"]) and np.array_equal(a["verts2d"], b["verts2d"])
    assert a["gt"]["verts_hand"].shape == (3, 778, 3) and a["K_roi_obj"].shape == (3, 3, 3)
    inits = synth.make_inits(a, 4, seed=7)
    assert inits["obj_R"].shape == (4, 3, 3, 3) and inits["pca"].shape == (4, 3, 16)
    # rotations stay rotations under the perturbation
    R = inits["obj_R"].reshape(-1, 3, 3).astype(np.float64)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3)[None], atol=1e-5)
    inp = synth.reference_inputs(dict(a, target_masks_object=np.zeros((3, 256, 256), np.float32),
                                      target_masks_hand=np.zeros((3, 256, 256), np.float32)), inits, 1)
    assert len(inp["person_parameters"]) == 3 and inp["person_parameters"][0]["mano_pca_pose"].shape == (1, 16)
    assert inp["object_parameters"][0]["K_roi"].shape == (1, 1, 3, 3)


def test_mano_asset_topology(mano_assets):
    a = mano_assets["right"]
    assert a["v_template"].shape == (778, 3) and a["f"].shape == (1538, 3) and a["closed_faces"].shape == (1552, 3)
    # the closed mesh is watertight: every edge is shared by exactly two faces
    f = a["closed_faces"]
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert (counts == 2).all()
    assert np.allclose(a["weights"].sum(1), 1, atol=1e-5) and np.allclose(a["J_regressor"].sum(1), 1, atol=1e-5)


def test_mano_blob_layout_matches_header(mano_assets):
    import re
    text = open(_lib.HEADER_PATH).read()
    blob = mano_blob(mano_assets["right"], 16, "cpu")
    assert blob.numel() == 778 * 3 + 778 * 30 + 135 * 2334 + 48 + 480 + 778 * 16 + 48 + 16 * 45
    assert re.search(r"#define HM_MANO_OFF_COMPS \(HM_MANO_OFF_MEAN \+ 48\)", text)
    off_w = 778 * 3 + 778 * 30 + 135 * 2334 + 48 + 480
    assert np.allclose(blob[off_w:off_w + 16].numpy(), mano_assets["right"]["weights"][0])
    off_j = 778 * 3 + 778 * 30 + 135 * 2334
    J = mano_assets["right"]["J_regressor"].astype(np.float64) @ mano_assets["right"]["v_template"].astype(np.float64)
    assert np.allclose(blob[off_j:off_j + 48].numpy().reshape(16, 3), J, atol=1e-6)


def test_header_signatures():
    sigs = _lib.parse_header()
    assert sigs["hm_raster_setup"] == "ppiiiiiiippp"
    assert sigs["hm_adam_step"] == "pppppifffpp"
    assert all(set(v) <= set("pif") for v in sigs.values())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_no_cpu_fallback(mano_assets):
    clip = synth.make_clip(2, "ellipsoid80", seed=1, mano_asset=mano_assets["right"])
    clip["target_masks_object"] = np.zeros((2, 256, 256), np.float32)
    clip["target_masks_hand"] = np.zeros((2, 256, 256), np.float32)
    batch = synth.make_batch(clip, synth.make_inits(clip, 1, seed=1))
    with pytest.raises(_lib.HomanB200Error):
        FitEngine(batch, synth.step1_loss_weights(), mano_asset=mano_assets["right"])
    from homan_b200 import ops
    with pytest.raises(_lib.HomanB200Error):
        ops.project(torch.zeros(1, 4, 3), torch.eye(3)[None])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_no_cpu_fallback_pose_fitter_and_render():
    """The rows added after the hot path (object-pose initialiser, RGB / depth render) fail loudly without a GPU too."""
    from homan_b200 import ops, pose_optimization as po
    verts, faces = synth.make_object("cube")
    with pytest.raises(_lib.HomanB200Error):
        po.PoseFitEngine(verts, faces, np.zeros((256, 256), np.float32), np.eye(3, dtype=np.float32),
                         np.zeros((2, 3, 2), np.float32), np.zeros((2, 3), np.float32))
    with pytest.raises(_lib.HomanB200Error):
        ops.render_rgbd(torch.zeros(1, 8, 3), torch.zeros(1, 12, 3, dtype=torch.int32), torch.zeros(1, 24, 3), 64)
    with pytest.raises(_lib.HomanB200Error):
        ops.rasterize_silhouettes(torch.zeros(1, 8, 3), torch.zeros(1, 12, 3, dtype=torch.int32), 64)


def test_header_declares_every_new_entry_point_with_a_reference_citation():
    text = open(_lib.HEADER_PATH).read()
    for name, cite in (("hm_offscreen_loss_fwd_bwd", "pose_optimization.py:112-134"),
                       ("hm_track_best", "pose_optimization.py:349-353"), ("hm_raster_shade", "homan/homan.py:510-613"),
                       ("hm_face_lighting", "Renderer.render"), ("hm_sdf_pair", "homan/eval/pointmetrics.py:102-124"),
                       ("hm_nearest_point", "homan/eval/pointmetrics.py:17-45")):
        assert name in _lib.SIGNATURES, name
        assert cite in text, cite
    handle = _lib.lib()
    assert all(hasattr(handle, n) for n in ("hm_offscreen_loss_fwd_bwd", "hm_track_best", "hm_raster_shade", "hm_face_lighting"))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_no_cpu_fallback_eval_metrics():
    from homan_b200.eval.pointmetrics import get_inter_metrics
    with pytest.raises(_lib.HomanB200Error):
        get_inter_metrics(torch.zeros(1, 778, 3), torch.zeros(1, 8, 3), torch.zeros(1, 12, 3, dtype=torch.int64),
                          torch.zeros(1, 12, 3, dtype=torch.int64))
    from homan_b200.eval.pointmetrics import nearest_dist2
    with pytest.raises(_lib.HomanB200Error):
        nearest_dist2(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_bench_prints_exactly_one_json_line_on_stdout(monkeypatch, capfd):
    """The driver parses bench.py's stdout as one JSON line: whatever a library writes to file descriptor 1 while the
    job runs (NCCL prints its version banner there when NCCL_DEBUG is set) must end up on stderr."""
    import importlib
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")

    def fake_run(args):
        os.write(1, b"NCCL version 0.0.0+test\n")   # a native library writing to fd 1
        print("a print from a helper")
        bench._LINE.append(json.dumps({"metric": "m", "value": 1.0}))

    monkeypatch.setattr(bench, "run_ours", fake_run)
    monkeypatch.setattr(bench, "_LINE", [])
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3"])
    bench.main()
    out, err = capfd.readouterr()
    assert out.strip().splitlines() == ['{"metric": "m", "value": 1.0}']
    assert "NCCL version" in err and "a print from a helper" in err
