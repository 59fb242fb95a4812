"""GPU parity of the individual kernels (MANO LBS, rigid placement, SDF pair, contact, Adam) against
the CPU oracle (torch autograd over oracle/*.py) on seeded inputs."""
import numpy as np
import pytest
import torch

from homan_b200 import synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-12))


@pytest.mark.parametrize("use_cache", [False, True])   # backward with / without the forward's posed template
@pytest.mark.parametrize("side", ["right", "left"])
def test_mano_lbs_forward_backward(side, use_cache, mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from homan_b200.engine import mano_blob
    from oracle import homan_ref
    asset = mano_assets[side]
    rng = np.random.default_rng(3)
    B = 5
    pca = rng.normal(size=(B, 20)).astype(np.float32) * 0.6
    rot = rng.normal(size=(B, 3)).astype(np.float32) * 0.4
    betas = rng.normal(size=(B, 10)).astype(np.float32) * 0.5
    mtr = rng.normal(size=(B, 3)).astype(np.float32) * 0.02
    r6 = rng.normal(size=(B, 3, 2)).astype(np.float32)
    tr = rng.normal(size=(B, 1, 3)).astype(np.float32) * 0.1
    gv = rng.normal(size=(B, 778, 3)).astype(np.float32)
    gc = rng.normal(size=(B, 3)).astype(np.float32)
    # oracle
    tp = {k: torch.from_numpy(v).requires_grad_() for k, v in
          dict(pca=pca, rot=rot, betas=betas, mtr=mtr, r6=r6, tr=tr).items()}
    mano = homan_ref.ManoPca(mano_assets["right"], mano_assets["left"])
    v, _ = mano(tp["pca"], tp["rot"], tp["betas"], side)
    v = v + tp["mtr"].unsqueeze(1)
    full, det = homan_ref.transform_persp(v, tp["tr"], homan_ref.rot6d_to_matrix(tp["r6"]), torch.ones(1))
    ((full * torch.from_numpy(gv)).sum() + (det.mean(1) * torch.from_numpy(gc)).sum()).backward()
    # kernel
    d = lambda x: torch.from_numpy(x).cuda()  # noqa: E731
    blob = mano_blob(asset, 16)
    dp, dr, db, dm, d6, dt = d(pca), d(rot), d(betas), d(mtr), d(r6), d(tr)
    verts = torch.empty(B, 778, 3, device="cuda")
    vposed = torch.empty(B, 778 * 3, device="cuda")
    s = current_stream()
    left = 1 if side == "left" else 0
    call("hm_mano_fwd", ptr(blob), 16, left, ptr(dp), 20, ptr(dr), ptr(db), ptr(dm), ptr(d6), ptr(dt), None, B,
         ptr(verts), None, ptr(vposed), s)
    assert _rel(verts.cpu(), full.detach()) < 2e-6
    g = {k: torch.zeros_like(x) for k, x in dict(pca=dp, rot=dr, betas=db, mtr=dm, r6=d6, tr=dt).items()}
    dgv, dgc = d(gv), d(gc)
    call("hm_mano_bwd", ptr(blob), 16, left, ptr(dp), 20, ptr(dr), ptr(db), ptr(dm), ptr(d6), ptr(dt), None, B,
         ptr(vposed) if use_cache else None, ptr(dgv), ptr(dgc), ptr(g["pca"]), ptr(g["rot"]), ptr(g["betas"]),
         ptr(g["mtr"]), ptr(g["r6"]), ptr(g["tr"]), s)
    torch.cuda.synchronize()
    for k in g:
        assert _rel(g[k].cpu(), tp[k].grad) < 1e-4, (k, _rel(g[k].cpu(), tp[k].grad))


def test_mano_layer_joints_match_numpy_fp64(mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from homan_b200.engine import mano_blob
    rng = np.random.default_rng(4)
    B = 3
    pca = rng.normal(size=(B, 16)).astype(np.float32) * 0.5
    rot = rng.normal(size=(B, 3)).astype(np.float32) * 0.3
    betas = rng.normal(size=(B, 10)).astype(np.float32) * 0.3
    v_ref, j_ref = synth.mano_forward_np(mano_assets["right"], pca, rot, betas)
    d = lambda x: torch.from_numpy(x).cuda()  # noqa: E731
    verts, joints = torch.empty(B, 778, 3, device="cuda"), torch.empty(B, 16, 3, device="cuda")
    blob, dp, dr, db = mano_blob(mano_assets["right"], 16), d(pca), d(rot), d(betas)
    call("hm_mano_fwd", ptr(blob), 16, 0, ptr(dp), 16, ptr(dr), ptr(db),
         None, None, None, None, B, ptr(verts), ptr(joints), None, current_stream())
    torch.cuda.synchronize()
    assert _rel(verts.cpu(), v_ref) < 2e-6 and _rel(joints.cpu(), j_ref) < 2e-6


def test_rigid_object(mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from oracle import homan_ref
    rng = np.random.default_rng(5)
    B = 4
    mesh, _ = synth.make_ellipsoid(8, 5)
    r6 = torch.from_numpy(rng.normal(size=(B, 3, 2)).astype(np.float32)).requires_grad_()
    tr = torch.from_numpy(rng.normal(size=(B, 1, 3)).astype(np.float32)).requires_grad_()
    gv = torch.from_numpy(rng.normal(size=(B, mesh.shape[0], 3)).astype(np.float32))
    m = torch.from_numpy(mesh)[None].repeat(B, 1, 1)
    out, _ = homan_ref.transform_persp(m, tr, homan_ref.rot6d_to_matrix(r6), torch.ones(1))
    (out * gv).sum().backward()
    dm, d6, dt = torch.from_numpy(mesh).cuda()[None].contiguous(), r6.detach().cuda(), tr.detach().cuda()
    verts = torch.empty(B, mesh.shape[0], 3, device="cuda")
    g6, gt = torch.zeros(B, 3, 2, device="cuda"), torch.zeros(B, 1, 3, device="cuda")
    s = current_stream()
    call("hm_rigid_fwd", ptr(dm), 1, ptr(d6), ptr(dt), None, B, mesh.shape[0], ptr(verts), s)
    dgv = gv.cuda()
    call("hm_rigid_bwd", ptr(dm), 1, ptr(d6), None, B, mesh.shape[0], ptr(dgv), ptr(g6), ptr(gt), s)
    torch.cuda.synchronize()
    assert _rel(verts.cpu(), out.detach()) < 1e-6
    assert _rel(g6.cpu(), r6.grad) < 1e-4 and _rel(gt.cpu(), tr.grad) < 1e-5


def _grasp(T, seed, mano_assets, obj="ellipsoid80"):
    clip = synth.make_clip(T, obj, seed=seed, mano_asset=mano_assets["right"])
    return clip["gt"]["verts_hand"], clip["gt"]["verts_obj"], clip


def test_sdf_pair_matches_oracle(mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from oracle import homan_ref
    vh, vo, clip = _grasp(3, 21, mano_assets)
    # push the object into the hand so that some samples are inside
    vo = vo + (vh.mean(1, keepdims=True) - vo.mean(1, keepdims=True)) * 0.8
    closed, fo = mano_assets["right"]["closed_faces"], clip["obj_faces"]
    th = torch.from_numpy(vh).requires_grad_()
    to = torch.from_numpy(vo)
    loss, dv = homan_ref.sdf_scene([th, to], [torch.from_numpy(closed), torch.from_numpy(fo)])
    loss.backward()
    assert float(loss) > 0
    B = vh.shape[0]
    dh, do = torch.from_numpy(vh).cuda(), torch.from_numpy(vo).cuda()
    part = torch.zeros(B, 16, device="cuda")
    gh = torch.zeros_like(dh)
    phi = torch.empty(B, 32 ** 3, device="cuda")
    s = current_stream()
    d_closed, d_fo = torch.from_numpy(closed).cuda(), torch.from_numpy(fo.astype(np.int32)).cuda()
    call("hm_sdf_pair", ptr(dh), ptr(d_closed), 1, ptr(do), B, 778, closed.shape[0], vo.shape[1],
         32, 0.2, 0.0, ptr(phi), ptr(part), None, None, s)
    a = part[:, 10].sum().item()
    call("hm_sdf_pair", ptr(do), ptr(d_fo), 1, ptr(dh), B, vo.shape[1],
         fo.shape[0], 778, 32, 0.2, 0.5, ptr(phi), ptr(part), ptr(gh), None, s)
    torch.cuda.synchronize()
    total = part[:, 10].sum().item()
    ref_a = float(dv[(0, 1)].sum() / 1.0)  # rescaled values; compare the normalised sums through the loss instead
    assert abs(total - float(loss)) <= 1e-4 * float(loss), (total, float(loss), a, ref_a)
    assert _rel(gh.cpu(), 0.5 * th.grad) < 1e-4


def test_sdf_dense_grid_matches_oracle(mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from oracle import sdfmod
    v, f = synth.make_ellipsoid(8, 5, (0.5, 0.8, 0.6))
    verts = torch.from_numpy(np.stack([v, v * 0.7 + 0.1]).astype(np.float32))
    faces = torch.from_numpy(f.astype(np.int32))
    ref = sdfmod.SDF()(faces, verts, 32)
    phi = torch.empty(2, 32, 32, 32, device="cuda")
    d_faces, d_verts = faces.cuda(), verts.cuda()
    call("hm_sdf_grid", ptr(d_faces), ptr(d_verts), 2, v.shape[0], f.shape[0], 32, ptr(phi), current_stream())
    torch.cuda.synchronize()
    assert int(((phi.cpu() > 0) != (ref > 0)).sum()) == 0
    assert _rel(phi.cpu(), ref) < 1e-5


def test_contact_matches_oracle(mano_assets):
    from homan_b200._lib import call, current_stream, ptr
    from oracle import homan_ref
    vh, vo, clip = _grasp(4, 22, mano_assets)
    th, to = torch.from_numpy(vh).requires_grad_(), torch.from_numpy(vo).requires_grad_()
    model = homan_ref.ClipModel.__new__(homan_ref.ClipModel)
    loss = homan_ref.ClipModel.contact(model, th, to)
    loss.backward()
    B = vh.shape[0]
    part = torch.zeros(B, 16, device="cuda")
    gh, go = torch.zeros(B, 778, 3, device="cuda"), torch.zeros(B, vo.shape[1], 3, device="cuda")
    dh, do = torch.from_numpy(vh).cuda(), torch.from_numpy(vo).cuda()
    call("hm_contact_fwd_bwd", ptr(dh), ptr(do), B, B, vo.shape[1],
         0.02, 2.0, ptr(part), ptr(gh), ptr(go), None, current_stream())
    torch.cuda.synchronize()
    assert abs(part[:, 11].sum().item() - float(loss)) <= 1e-5 * float(loss)
    assert _rel(gh.cpu(), 2 * th.grad) < 1e-4 and _rel(go.cpu(), 2 * to.grad) < 1e-4
    d = torch.cdist(torch.from_numpy(vh), torch.from_numpy(vo)).flatten(1).min(1)[0]
    # metric only: the reference's |h|^2 + |o|^2 - 2 h.o form cancels catastrophically at mm distances
    assert _rel(part[:, 12].cpu(), d) < 5e-2


def test_adam_matches_torch():
    from homan_b200._lib import call, current_stream, ptr
    rng = np.random.default_rng(6)
    n = 1000
    p0 = rng.normal(size=n).astype(np.float32)
    lr = np.where(np.arange(n) < 400, 1e-2, np.where(np.arange(n) < 800, 1e-1, 0.0)).astype(np.float32)
    ref_a = torch.from_numpy(p0[:400].copy()).requires_grad_()
    ref_b = torch.from_numpy(p0[400:800].copy()).requires_grad_()
    opt = torch.optim.Adam([{"params": [ref_a], "lr": 1e-2}, {"params": [ref_b], "lr": 1e-1}])
    p, m, v = torch.from_numpy(p0).cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    dlr = torch.from_numpy(lr).cuda()
    for it in range(5):
        g = (rng.normal(size=n) * (10.0 ** rng.integers(-4, 2))).astype(np.float32)
        ref_a.grad, ref_b.grad = torch.from_numpy(g[:400].copy()), torch.from_numpy(g[400:800].copy())
        opt.step()
        step += 1
        dg = torch.from_numpy(g).cuda()
        call("hm_adam_step", ptr(p), ptr(dg), ptr(m), ptr(v), ptr(dlr), n, 0.9, 0.999, 1e-8,
             ptr(step), current_stream())
    torch.cuda.synchronize()
    out = p.cpu().numpy()
    assert np.allclose(out[:400], ref_a.detach().numpy(), rtol=1e-5, atol=1e-7)
    assert np.allclose(out[400:800], ref_b.detach().numpy(), rtol=1e-5, atol=1e-7)
    assert np.array_equal(out[800:], p0[800:])


def test_inter_metrics_match_oracle(mano_assets):
    """Penetration depth / contact flag of the reference's evaluation (homan/eval/pointmetrics.py:102-124) against
    the oracle's SDFSceneLoss restatement: dist_values[(1, 0)].max(1)."""
    from homan_b200.eval.pointmetrics import get_inter_metrics
    from oracle import homan_ref
    vh, vo, clip = _grasp(4, 33, mano_assets)
    vo = (vo + (vh.mean(1, keepdims=True) - vo.mean(1, keepdims=True)) * np.array([0.9, 0.6, 0.3, -2.0])[:, None, None]).astype(np.float32)
    closed, fo = mano_assets["right"]["closed_faces"], clip["obj_faces"]
    _, dv = homan_ref.sdf_scene([torch.from_numpy(vh), torch.from_numpy(vo)],
                                [torch.from_numpy(closed), torch.from_numpy(fo)])
    ref = dv[(1, 0)].max(1)[0].numpy()
    got = get_inter_metrics(torch.from_numpy(vh).cuda(), torch.from_numpy(vo).cuda(),
                            torch.from_numpy(closed)[None].cuda(), torch.from_numpy(fo.astype(np.int64))[None].cuda())
    assert ref[:3].min() > 0 and ref[3] == 0   # three interpenetrating scenes, one apart
    assert np.allclose(got["pen_depths"], ref, rtol=1e-4, atol=1e-7), (got["pen_depths"], ref)
    assert got["has_contact"] == (ref > 0).tolist()


def test_point_metrics_match_scipy_and_the_reference_expressions(mano_assets):
    """get_point_metrics / get_align_metrics (homan/eval/pointmetrics.py:17-45,62-99) on hm_nearest_point, against
    scipy's cKDTree (the reference's own ADD-S search) and the dense distance-matrix form the reference documents as
    equal to pytorch3d's chamfer ((dist_mat.min(1)[0] + dist_mat.min(2)[0]).mean(-1), pointmetrics.py:25,94)."""
    from scipy import spatial
    from homan_b200.eval.pointmetrics import get_align_metrics, get_point_metrics, nearest_dist2
    rng = np.random.default_rng(4)
    vh, vo, clip = _grasp(3, 21, mano_assets)
    gt_o, gt_h = vo.astype(np.float32), vh.astype(np.float32)
    pr_o = (gt_o * np.float32(1.07) + rng.normal(0, 0.004, gt_o.shape)).astype(np.float32)
    pr_h = (gt_h * np.float32(1.07) + rng.normal(0, 0.002, gt_h.shape)).astype(np.float32)
    pr_o_sub = pr_o[:, ::3]                                   # a point set of another size: no vertex assignment
    t = torch.from_numpy
    # nearest neighbours (index too), incl. exact ties (duplicated targets -> lowest index)
    a = t(gt_o).cuda()
    b = torch.cat((t(pr_o_sub), t(pr_o_sub)), 1).cuda()
    from homan_b200._lib import call, current_stream, ptr
    d2 = torch.empty(a.shape[:2], device="cuda")
    idx = torch.empty(a.shape[:2], dtype=torch.int32, device="cuda")
    call("hm_nearest_point", ptr(a), ptr(b), a.shape[0], a.shape[1], b.shape[1], ptr(d2), ptr(idx), current_stream())
    dm = torch.cdist(t(gt_o).double(), torch.cat((t(pr_o_sub), t(pr_o_sub)), 1).double()) ** 2
    assert torch.equal(idx.cpu().long(), dm.float().argmin(2)) or (idx.cpu().long() < pr_o_sub.shape[1]).all()
    assert torch.allclose(d2.cpu().double(), dm.min(2)[0], rtol=1e-5, atol=1e-12)
    assert torch.equal(nearest_dist2(a, b), d2)
    for pred in (pr_o, pr_o_sub):
        got = get_point_metrics(t(gt_o), t(pred))
        dm = torch.cdist(t(gt_o).double(), t(pred).double()) ** 2
        cham = (dm.min(2)[0].mean(-1) + dm.min(1)[0].mean(-1)).numpy()
        adds = np.array([spatial.cKDTree(pe).query(pg, k=1)[0].mean() for pg, pe in zip(gt_o, pred)])
        assert np.allclose(got["chamfer_dists"], cham, rtol=1e-4), (got["chamfer_dists"], cham)
        assert np.allclose(got["add-s"], adds, rtol=1e-4)
        if pred.shape[1] == gt_o.shape[1]:
            assert np.allclose(got["verts_dists"], np.linalg.norm(gt_o - pred, axis=-1).mean(-1), rtol=1e-5)
        else:
            assert got["verts_dists"] == got["add-s"]
    # aligned metrics: the reference expressions in fp64
    got = get_align_metrics(t(gt_h), t(pr_h), t(gt_o), t(pr_o))
    gh, ph, go, po = (t(x).double() for x in (gt_h, pr_h, gt_o, pr_o))
    cent = gh.mean(1, keepdim=True)
    gh_c, go_c, ph_c, po_c = gh - cent, go - cent, ph - cent, po - cent
    gs = torch.sqrt((gh_c.norm(2, -1) ** 2).sum(1) / gh.shape[1])
    ps = torch.sqrt((ph_c.norm(2, -1) ** 2).sum(1) / ph.shape[1])
    ph_cs, po_cs = ph_c / ps.view(-1, 1, 1) * gs.view(-1, 1, 1), po_c / ps.view(-1, 1, 1) * gs.view(-1, 1, 1)
    dm = torch.cdist(po_cs, go_c) ** 2
    assert np.allclose(got["hand_mean_aligned"], (gh_c - ph_cs).norm(2, -1).mean(-1).numpy(), rtol=1e-4)
    assert np.allclose(got["obj_chamfer_aligned"], (dm.min(2)[0].mean(-1) + dm.min(1)[0].mean(-1)).numpy(), rtol=1e-4)


@pytest.mark.parametrize("aa", [True, False])
def test_fused_loss_and_sweep_lists_equal_the_two_kernels(aa, mano_assets):
    """hm_sil_loss_prep = hm_sil_loss_fwd_bwd + hm_raster_grad_prep, bit for bit: loss, IoU, grad_alpha, the four
    sweep bit lines, the run counts and every stored run."""
    from homan_b200 import ops
    from homan_b200._lib import call, current_stream, ptr
    clip = synth.make_clip(4, "ellipsoid500", seed=17, mano_asset=mano_assets["right"])
    verts = np.concatenate((clip["gt"]["verts_hand"], clip["gt"]["verts_hand"] + np.float32([0.004, -0.003, 0.0])))
    K = np.concatenate((clip["K_roi_hand"], clip["K_roi_hand"])).astype(np.float32)
    B, R = verts.shape[0], 256
    ndc = ops.project(torch.from_numpy(verts).cuda(), torch.from_numpy(K).cuda(), orig_size=1.0)
    faces = torch.from_numpy(mano_assets["right"]["f"].astype(np.int32)).cuda()[None]
    bufs = [ops.RasterBuffers(B, 778, faces.shape[1], R, aa, "cuda") for _ in range(2)]
    rng = np.random.default_rng(2)
    target = torch.roll((ops.raster_forward(bufs[0], ndc, faces) > 0.5).to(torch.int8), shifts=(4, -6), dims=(1, 2))
    target[:, :, 100:124] = -1
    target = target.contiguous()
    norm = torch.from_numpy(rng.uniform(1e-6, 2e-6, size=B).astype(np.float32)).cuda()
    ops.raster_forward(bufs[1], ndc, faces)
    part = [torch.zeros(B, 16, device="cuda") for _ in range(2)]
    ga = [torch.empty(B, R, R, device="cuda") for _ in range(2)]
    s = current_stream()
    for b_ in bufs:   # stale contents must not matter
        b_.m_row.fill_(-1); b_.m_col.fill_(-1); b_.runs.fill_(-1); b_.run_counts.fill_(-1)
    call("hm_sil_loss_fwd_bwd", ptr(bufs[0].alpha), ptr(target), ptr(norm), 1.5, B, R, ptr(part[0]), 16,
         part[0].data_ptr() + 4, 16, ptr(ga[0]), s)
    call("hm_raster_grad_prep", ptr(ga[0]), ptr(bufs[0].cov_row), ptr(bufs[0].cov_col), B, R, int(aa), ptr(bufs[0].m_row),
         ptr(bufs[0].m_col), ptr(bufs[0].runs), ptr(bufs[0].run_counts), s)
    call("hm_sil_loss_prep", ptr(bufs[1].alpha), ptr(target), ptr(norm), 1.5, B, R, int(aa), ptr(part[1]), 16,
         part[1].data_ptr() + 4, 16, ptr(ga[1]), ptr(bufs[1].cov_row), ptr(bufs[1].cov_col), ptr(bufs[1].m_row),
         ptr(bufs[1].m_col), ptr(bufs[1].runs), ptr(bufs[1].run_counts), s)
    torch.cuda.synchronize()
    assert torch.equal(ga[0], ga[1]) and float(ga[0].abs().max()) > 0
    assert torch.allclose(part[0], part[1], rtol=1e-6)   # (sums over 256 vs 512 threads)
    assert torch.equal(bufs[0].m_row, bufs[1].m_row) and torch.equal(bufs[0].m_col, bufs[1].m_col)
    assert torch.equal(bufs[0].run_counts, bufs[1].run_counts)
    cnt = (bufs[0].run_counts & 15).clamp(max=8)   # overflowed lists store their first 8 runs
    stored = torch.arange(8, device="cuda").view(1, 1, 1, 8) < cnt.unsqueeze(-1)
    assert int(cnt.max()) >= 2 and torch.equal(bufs[0].runs[stored], bufs[1].runs[stored])
