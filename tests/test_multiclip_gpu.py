"""GPU: multi-clip batches (BASELINE.json config 4: whole clips per GPU, every clip with its own object mesh, as the
reference fits a different object per sample - fit_vid_dataset.py:190-296). A clip inside a multi-clip batch follows
the trajectory it has when fitted alone, the per-clip argmin picks the best init of every clip, and the batch agrees
with the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clip_slice(batch, c):
    sel = np.asarray(batch["clip_of_problem"]) == c
    out = {}
    for k, v in batch.items():
        if isinstance(v, np.ndarray) and v.shape[:1] == (batch["P"],) and k not in ("obj_verts_can", "obj_faces", "hand_faces", "clip_ids"):
            v = v[sel]
        out[k] = v
    out["obj_verts_can"], out["obj_faces"] = batch["obj_verts_can"][c], batch["obj_faces"][c]
    out.pop("clip_of_problem")
    out["P"] = int(sel.sum())
    return out


def test_clips_with_their_own_objects(mano_assets):
    from homan_b200.engine import FitEngine
    from homan_b200.workload import CONFIGS, make_workload
    from oracle import homan_ref
    asset = mano_assets["right"]
    batch, lw = make_workload("tiny4", mano_asset=asset)
    C, inits = CONFIGS["tiny4"]["clips"], CONFIGS["tiny4"]["P"]
    assert batch["obj_verts_can"].shape[0] == C and batch["P"] == C * inits
    assert not np.allclose(batch["obj_verts_can"][0], batch["obj_verts_can"][1])   # different objects
    full = FitEngine(batch, lw, mano_asset=asset, use_graph=True)
    out = full.fit(3)
    # (1) the oracle on every problem with its clip's mesh: first-iteration losses
    ref = homan_ref.evaluate(batch, lw, mano_assets={"right": asset})
    first = {k: v[0] for k, v in out["losses"].items()}
    for k, v in ref["losses"].items():
        if k in first:
            tol = 1e-3 if "sil" in k else 1e-4
            assert np.all(np.abs(first[k] - v[0]) <= tol * np.maximum(np.abs(v[0]), 1e-7) + 1e-9), (k, first[k], v[0])
    # (2) a clip fitted alone follows the same first iterations
    for c in range(C):
        alone = FitEngine(_clip_slice(batch, c), lw, mano_asset=asset, use_graph=False).fit(2)
        assert np.allclose(alone["total"], out["total"][:2, c * inits:(c + 1) * inits], rtol=1e-4), c
    # (3) per-clip argmin over the inits
    bi, bl = full.best_init(clips=C)
    tot = out["total"][-1].reshape(C, inits)
    assert np.array_equal(bi.cpu().numpy(), tot.argmin(1)) and np.allclose(bl.cpu().numpy(), tot.min(1))
