"""Generates tests/golden/ref_state_dict.npz: the checkpoint schema of the reference (`joint_fit.pt`,
/root/reference/fit_vid_dataset.py:365-372: HOMan.state_dict() minus the `mano_model.*` entries) recorded from the
UNMODIFIED reference after a short fit on CPU (oracle/refshim.py), plus the losses the reference evaluates at that
state - SURVEY.md 8f row 2 (checkpoint / wire formats). Build container only.

    python scripts/make_golden_state.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_utils import load, reference_inputs  # noqa: E402
from homan_b200 import synth  # noqa: E402
from oracle import refshim  # noqa: E402


def main():
    assert refshim.reference_available(), "needs /root/reference"
    assets = {"right": synth.make_mano_asset(0, "right"), "left": synth.make_mano_asset(1, "left")}
    scratch = tempfile.mkdtemp(prefix="homan_golden_state_")
    refshim.install(scratch, assets)
    z, batch, lw, _ = load("ref_small_step2", assets["right"])
    inp = reference_inputs(batch, 0, assets["right"])
    model, ev = refshim.run_reference_fit(inp, lw, 2, scratch, lr=1e-2)
    sd = model.state_dict()
    out = {"all_keys": np.array(sorted(sd.keys())),
           "all_shapes": np.array([",".join(map(str, sd[k].shape)) for k in sorted(sd.keys())])}
    saved = {k: v for k, v in sd.items() if "mano_model" not in k}   # fit_vid_dataset.py:365-372
    for k, v in saved.items():
        a = v.numpy()
        out["sd_" + k] = a.astype(np.int8) if ("mask" in k and a.dtype != np.bool_ and set(np.unique(a)) <= {0.0, 1.0}) else a
    with torch.no_grad():
        pass
    loss_dict, metric_dict = model(lw)      # the losses at the checkpointed state
    for k, v in loss_dict.items():
        out["eval_" + k] = np.asarray(float(v))
    for k, v in metric_dict.items():
        out["metric_" + k] = np.asarray(float(v))
    path = os.path.join(ROOT, "tests", "golden", "ref_state_dict.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")
    for k in sorted(saved):
        print("  ", k, tuple(saved[k].shape), saved[k].dtype)
    print({k: float(v) for k, v in loss_dict.items()})


if __name__ == "__main__":
    main()
