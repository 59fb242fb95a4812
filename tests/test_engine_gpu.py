"""GPU parity of the fused fitting engine against golden vectors produced by the UNMODIFIED reference
Python (scripts/make_golden.py -> tests/golden/*.npz) and against the CPU oracle.

Tolerances: loss scalars 1e-4 relative (north_star); parameter gradients 1e-4 of the gradient scale,
except silhouette-driven ones where a one-ulp difference in the projected vertices may move a face
boundary across a pixel centre (1e-3)."""
import numpy as np
import pytest
import torch

from golden_utils import PARAMS, load

pytestmark = pytest.mark.gpu
CASES = ["ref_cfg1_cube", "ref_small_step1", "ref_small_step2"]


def _engine(name, mano_assets, **kw):
    from homan_b200.engine import FitEngine
    z, batch, lw, iters = load(name, mano_assets["right"])
    return z, batch, lw, iters, FitEngine(batch, lw, lr=1e-2, mano_asset=mano_assets["right"], **kw)


@pytest.mark.parametrize("name", CASES)
def test_first_iteration_matches_reference(name, mano_assets):
    z, batch, lw, _, eng = _engine(name, mano_assets, use_graph=False)
    losses = eng.evaluate()
    torch.cuda.synchronize()
    total = eng.total.cpu().numpy()
    report = []
    for p in range(batch["P"]):
        for k, v in losses.items():
            ref = z[f"ev_{k}_p{p}"][0]
            ok = abs(v[p] - ref) <= 1e-4 * max(abs(ref), 1e-7) + 1e-9
            report.append((k, p, float(v[p]), float(ref), ok))
        ref = z[f"ev_loss_p{p}"][0]
        report.append(("total", p, float(total[p]), float(ref), abs(total[p] - ref) <= 1e-4 * abs(ref)))
    bad = [r for r in report if not r[-1]]
    assert not bad, bad
    T = batch["T"]
    gbad = []
    for p in range(batch["P"]):
        for k in PARAMS:
            key = f"grad0_{k}_p{p}"
            if key not in z.files:
                continue
            g_ref = z[key]
            g = eng.grads[k].view(batch["P"], T, *eng.grads[k].shape[1:])[p].cpu().numpy().reshape(g_ref.shape)
            scale = np.abs(g_ref).max()
            tol = 1e-3 if "object" in k else 1e-4
            err = np.abs(g - g_ref).max()
            if not err <= tol * scale + 1e-10:
                gbad.append((k, p, float(err), float(scale)))
    assert not gbad, gbad


@pytest.mark.parametrize("name", CASES)
def test_trajectory_matches_reference(name, mano_assets):
    """A few Adam steps. Iterations 0-1 must agree to 1e-4 (the first Adam step is sign-like, so it is
    insensitive to gradient noise). Afterwards the trajectory is only boundedly reproducible: coverage is a
    discontinuous function of the pose, so fp32-level differences (atomic summation order, one-ulp projection
    differences) flip boundary pixels and Adam's normalised update amplifies that (the reference's own
    neural_renderer backward is order-nondeterministic in the same way)."""
    z, batch, lw, iters, eng = _engine(name, mano_assets, use_graph=True)
    out = eng.fit(iters)
    for p in range(batch["P"]):
        ref = z[f"ev_loss_p{p}"]
        got = out["total"][:, p]
        assert abs(got[0] - ref[0]) <= 1e-4 * abs(ref[0]), (got, ref)
        assert np.all(np.abs(got[:2] - ref[:2]) <= 1e-4 * np.abs(ref[:2])), (got, ref)
        k = min(6, len(ref))   # beyond a handful of steps the comparison is chaotic (see docstring); the
        # teacher-forced test below covers every iteration of the reference trajectory
        assert np.all(np.abs(got[:k] - ref[:k]) <= 5e-2 * np.abs(ref[:k])), (got, ref)
        for name in ("translations_object", "translations_hand"):
            T = batch["T"]
            fin = out["params"][name].reshape(batch["P"], T, 1, 3)[p]
            assert np.abs(fin - z[f"final_{name}_p{p}"].reshape(T, 1, 3)).max() < 2e-2


@pytest.mark.parametrize("name", CASES)
def test_losses_along_the_reference_trajectory(name, mano_assets):
    """Teacher-forced parity: at the parameters the unmodified reference visited at every iteration
    (traj_* in the golden file) the engine's loss scalars equal the reference's (1e-4 relative; 1e-3 on the
    silhouette term, where a one-ulp difference in a projected vertex may flip a boundary sub-pixel)."""
    z, batch, lw, iters, eng = _engine(name, mano_assets, use_graph=False)
    P, T = batch["P"], batch["T"]
    worst = 0.0
    for it in range(iters):
        for k in ("translations_object", "rotations_object", "translations_hand", "rotations_hand",
                  "mano_pca_pose", "mano_betas"):
            val = np.stack([z[f"traj_{k}_p{p}"][it] for p in range(P)])
            eng.params[k].copy_(torch.from_numpy(val).reshape(eng.params[k].shape))
        losses = eng.evaluate()
        total = eng.total.cpu().numpy()
        for p in range(P):
            for k, v in losses.items():
                ref = z[f"ev_{k}_p{p}"][it]
                tol = 1e-3 if "sil" in k else 1e-4
                assert abs(v[p] - ref) <= tol * max(abs(ref), 1e-7) + 1e-9, (it, k, p, v[p], ref)
            ref = z[f"ev_loss_p{p}"][it]
            worst = max(worst, abs(total[p] - ref) / abs(ref))
            assert abs(total[p] - ref) <= 1e-3 * abs(ref), (it, p, total[p], ref)
    print("worst relative total-loss error along the trajectory:", worst)


def test_graph_replay_equals_eager(mano_assets):
    _, _, _, _, e1 = _engine("ref_small_step2", mano_assets, use_graph=False)
    _, _, _, _, e2 = _engine("ref_small_step2", mano_assets, use_graph=True)
    o1, o2 = e1.fit(3), e2.fit(3)
    # atomics make the raster / contact gradient sums order dependent: the first two iterations agree to
    # fp32 noise, later ones up to the discrete events (boundary pixel flips) that noise can trigger
    assert np.allclose(o1["total"][:2], o2["total"][:2], rtol=1e-5)
    assert np.allclose(o1["total"], o2["total"], rtol=5e-2)
    for k in o1["params"]:
        assert np.allclose(o1["params"][k], o2["params"][k], rtol=5e-2, atol=5e-2), k


def test_problems_are_independent(mano_assets):
    """Problem p of a batch follows the trajectory it has when run alone (per-problem normalisers)."""
    from homan_b200.engine import FitEngine
    z, batch, lw, _ = load("ref_small_step2", mano_assets["right"])
    full = FitEngine(batch, lw, mano_asset=mano_assets["right"], use_graph=False).fit(2)
    sub = {k: (v[1:2] if isinstance(v, np.ndarray) and v.shape[:1] == (batch["P"],) and k not in
               ("obj_verts_can", "obj_faces", "hand_faces") else v) for k, v in batch.items()}
    sub["P"] = 1
    one = FitEngine(sub, lw, mano_asset=mano_assets["right"], use_graph=False).fit(2)
    assert np.allclose(full["total"][:, 1], one["total"][:, 0], rtol=1e-5)


def test_argmin_over_inits(mano_assets):
    _, batch, _, _, eng = _engine("ref_small_step2", mano_assets, use_graph=False)
    eng.fit(1)
    bi, bl = eng.best_init(clips=1)
    tot = eng.total.cpu().numpy()
    assert int(bi[0]) == int(np.argmin(tot)) and abs(float(bl[0]) - tot.min()) < 1e-7
