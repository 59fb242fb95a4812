"""Synthetic workload generator (pure numpy, seeded) for tests and bench.py.

The reference ships no dataset that can be used offline and MANO is licence-gated
(/root/reference/README.md:72-90), so every benchmark / parity input is synthetic:

* a MANO-shaped hand asset (778 vertices, 1538 open / 1552 watertight faces, 16 joints,
  10 betas, 135 pose-corrective features, 45x45 PCA basis) with the same array layout the
  real `MANO_RIGHT.pkl` has (`v_template, shapedirs, posedirs, J_regressor, weights,
  hands_components, hands_mean`), so that a real asset can be swapped in,
* closed ellipsoid / cube object meshes,
* per-frame dictionaries following the input schema of
  `optimize_hand_object` (/root/reference/homan/jointopt.py:22-124).
"""
import math

import numpy as np

MANO_PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], dtype=np.int32)
N_HAND_VERTS = 778
N_HAND_FACES = 1538
N_HAND_FACES_CLOSED = 1552


# ----------------------------------------------------------------------------- meshes
def make_cube(side=0.10):
    """8 vertices / 12 outward-facing triangles, centred."""
    h = side / 2.0
    v = np.array([[x, y, z] for x in (-h, h) for y in (-h, h) for z in (-h, h)], dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, c, d in quads:
        f += [(a, b, c), (a, c, d)]
    f = np.array(f, dtype=np.int32)
    # make outward facing (positive signed volume)
    vol = np.einsum("ij,ij->i", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])).sum()
    if vol < 0:
        f = f[:, ::-1].copy()
    return v, f


def _uv_sphere_topology(L, K):
    """Closed genus-0 triangulation: K rings of L vertices + 2 poles. V = L*K+2, F = 2*L*K.

    Vertex order: north pole, ring 0 .. ring K-1, south pole. Faces wind outward for a
    sphere whose pole axis is +z and whose longitudes grow counter-clockwise seen from +z.
    """
    def ring(j, i):
        return 1 + j * L + (i % L)

    north, south = 0, 1 + L * K
    faces = []
    for i in range(L):
        faces.append((north, ring(0, i), ring(0, i + 1)))
    for j in range(K - 1):
        for i in range(L):
            a, b, c, d = ring(j, i), ring(j + 1, i), ring(j + 1, i + 1), ring(j, i + 1)
            faces.append((a, b, c))
            faces.append((a, c, d))
    for i in range(L):
        faces.append((south, ring(K - 1, i + 1), ring(K - 1, i)))
    return np.array(faces, dtype=np.int32)


def make_ellipsoid(L=25, K=10, semi_axes=(0.04, 0.06, 0.04)):
    """Closed ellipsoid, F = 2*L*K faces, V = L*K + 2 vertices (500 f / 252 v by default)."""
    faces = _uv_sphere_topology(L, K)
    verts = [(0.0, 0.0, 1.0)]
    for j in range(K):
        th = math.pi * (j + 1) / (K + 1)
        for i in range(L):
            ph = 2 * math.pi * i / L
            verts.append((math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)))
    verts.append((0.0, 0.0, -1.0))
    verts = np.array(verts, dtype=np.float64) * np.array(semi_axes)[None]
    return verts.astype(np.float32), faces


# ----------------------------------------------------------------------------- hand asset
_FINGER_ANGLES = np.deg2rad([20.0, 0.0, -40.0, -20.0, 75.0])  # index, middle, pinky, ring, thumb
_FINGER_LEN = np.array([0.075, 0.082, 0.058, 0.074, 0.055])
_PALM_R = 0.042


def _hand_outline(phi):
    r = np.full_like(phi, _PALM_R)
    for a, ln in zip(_FINGER_ANGLES, _FINGER_LEN):
        d = np.angle(np.exp(1j * (phi - a)))
        r = r + ln * np.exp(-(d / 0.105) ** 2)
    return r


_TEMPLATE_CACHE = {}


def _hand_template(mesh):
    """(verts [778,3] float64, closed faces [1552,3]; the first 1538 are the open raster mesh) of the synthetic hand.

    "delaunay" (default): well-shaped faces, homan_b200/data/hand_template.npz made by scripts/make_hand_template.py
    (planar Delaunay pillow of the mitten outline; longest / shortest edge of a face: median 1.07, worst 3.4).
    "polar": the round-1 surface, a 97 x 8 polar triangulation of the same outline (many sliver faces: about three times
    the scan-line crossings per face of a well-shaped mesh); kept so that earlier measurements stay comparable."""
    if mesh in _TEMPLATE_CACHE:
        return _TEMPLATE_CACHE[mesh]
    if mesh == "delaunay":
        import os
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hand_template.npz"))
        verts, closed = z["verts"].astype(np.float64), z["closed_faces"].astype(np.int32)
    elif mesh == "polar":
        L, K = 97, 8
        closed = _uv_sphere_topology(L, K)
        # geometry: flat star-shaped pillow, pole axis = palm normal (z)
        verts = np.zeros((N_HAND_VERTS, 3), dtype=np.float64)
        thick = 0.011
        verts[0] = (0, 0, thick)
        for j in range(K):
            th = math.pi * (j + 1) / (K + 1)
            for i in range(L):
                ph = 2 * math.pi * i / L
                r = _hand_outline(np.array([ph]))[0] * math.sin(th) ** 0.6
                verts[1 + j * L + i] = (r * math.cos(ph), r * math.sin(ph), thick * math.cos(th))
        verts[-1] = (0, 0, -thick)
        # the 14 cap faces (a strip of 7 quads at the wrist side, phi ~ pi, between rings 3 and 4) go last
        cap = []
        i0 = L // 2 - 3
        base = L + 2 * L * 3  # faces before band j=3
        for i in range(i0, i0 + 7):
            cap += [base + 2 * i, base + 2 * i + 1]
        keep = np.ones(len(closed), dtype=bool)
        keep[cap] = False
        closed = np.concatenate([closed[keep], closed[~keep]], 0)
    else:
        raise ValueError(mesh)
    assert verts.shape == (N_HAND_VERTS, 3) and closed.shape == (N_HAND_FACES_CLOSED, 3)
    _TEMPLATE_CACHE[mesh] = (verts, closed)
    return verts, closed


def make_mano_asset(seed=0, side="right", mesh="delaunay"):
    """Seeded synthetic MANO-shaped asset. Returns a dict of float32 / int32 numpy arrays."""
    rng = np.random.default_rng(seed)
    verts, closed = _hand_template(mesh)
    verts, closed = verts.copy(), closed.copy()
    faces_open = closed[:N_HAND_FACES]
    # joints: wrist + 3 per finger (index, middle, pinky, ring, thumb)
    targets = [(-0.025, 0.0, 0.0)]
    for a, ln in zip(_FINGER_ANGLES, _FINGER_LEN):
        for s in (0.0, 0.36, 0.68):
            rr = _PALM_R * 0.95 + s * ln
            targets.append((rr * math.cos(a), rr * math.sin(a), 0.0))
    targets = np.array(targets)
    d2 = ((verts[None] - targets[:, None]) ** 2).sum(-1)  # [16, 778]
    jreg = np.exp(-d2 / (2 * 0.008 ** 2))
    jreg /= jreg.sum(1, keepdims=True)
    joints = jreg @ verts
    # skinning weights: distance to the bone segment joint -> child-ish end
    ends = joints.copy()
    for j in range(16):
        kids = np.where(MANO_PARENTS == j)[0]
        if j == 0:
            ends[j] = joints[j] + np.array([0.03, 0, 0])
        elif len(kids):
            ends[j] = joints[kids[0]]
        else:
            f = (j - 1) // 3
            a = _FINGER_ANGLES[f]
            ends[j] = joints[j] + 0.32 * _FINGER_LEN[f] * np.array([math.cos(a), math.sin(a), 0])
    wts = np.zeros((N_HAND_VERTS, 16))
    for j in range(16):
        ab = ends[j] - joints[j]
        tt = np.clip(((verts - joints[j]) @ ab) / (ab @ ab), 0, 1)
        closest = joints[j] + tt[:, None] * ab
        wts[:, j] = np.exp(-((verts - closest) ** 2).sum(-1) / (2 * 0.0075 ** 2))
    wts[:, 0] += 1e-6
    wts /= wts.sum(1, keepdims=True)
    # blend shapes: smooth low-frequency fields
    shapedirs = np.zeros((N_HAND_VERTS, 3, 10))
    for k in range(10):
        fr = rng.normal(size=(3, 3)) * 12.0
        ph = rng.uniform(0, 2 * math.pi, size=3)
        shapedirs[:, :, k] = 0.004 * np.sin(verts @ fr.T + ph[None]) + 0.03 * rng.normal() * verts
    posedirs = (rng.normal(size=(135, N_HAND_VERTS, 3)) * 4e-4)
    # localise each corrective to vertices skinned to that joint or its parent
    for j in range(15):
        infl = np.clip(wts[:, j + 1] + wts[:, MANO_PARENTS[j + 1]], 0, 1)
        posedirs[9 * j:9 * j + 9] *= infl[None, :, None]
    posedirs = posedirs.reshape(135, N_HAND_VERTS * 3)
    comps, _ = np.linalg.qr(rng.normal(size=(45, 45)))
    comps = comps * rng.uniform(0.4, 1.0, size=(45, 1))
    hands_mean = rng.normal(size=45) * 0.12
    if side == "left":
        verts[:, 0] *= -1
        shapedirs[:, 0, :] *= -1
        pd = posedirs.reshape(135, N_HAND_VERTS, 3)
        pd[:, :, 0] *= -1
        posedirs = pd.reshape(135, -1)
        closed = closed[:, ::-1].copy()
        faces_open = closed[:N_HAND_FACES]
        jreg = jreg  # unchanged: joints mirror with the vertices
    return {
        "v_template": verts.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "J_regressor": jreg.astype(np.float32),
        "weights": wts.astype(np.float32),
        "hands_components": comps.astype(np.float32),
        "hands_mean": hands_mean.astype(np.float32),
        "parents": MANO_PARENTS.copy(),
        "f": faces_open.astype(np.int32).copy(),
        "closed_faces": closed.astype(np.int32).copy(),
    }


# ----------------------------------------------------------------------------- numpy MANO forward (fp64)
def rodrigues_np(r):
    """[N,3] axis-angle -> [N,3,3] with the smplx convention angle = ||r + 1e-8||."""
    r = np.asarray(r, dtype=np.float64)
    angle = np.linalg.norm(r + 1e-8, axis=1, keepdims=True)
    d = r / angle
    c, s = np.cos(angle)[:, :, None], np.sin(angle)[:, :, None]
    Kx = np.zeros((r.shape[0], 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -d[:, 2], d[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = d[:, 2], -d[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -d[:, 1], d[:, 0]
    return np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)


def mano_forward_np(asset, pca_pose, rot, betas, side="right", ncomps=16):
    """fp64 numpy restatement of ManoModel.forward_pca (/root/reference/homan/manomodel.py:84-151)
    on a dict asset. pca_pose [B,>=ncomps], rot [B,3], betas [B,10] -> verts [B,778,3], joints [B,16,3]."""
    a = {k: np.asarray(v, dtype=np.float64) for k, v in asset.items() if k not in ("parents", "f", "closed_faces")}
    parents = np.asarray(asset["parents"])
    B = pca_pose.shape[0]
    hand_pose = np.asarray(pca_pose, np.float64)[:, :ncomps] @ a["hands_components"][:ncomps]
    if side == "left":
        hand_pose[:, 1::3] *= -1
        hand_pose[:, 2::3] *= -1
    hand_pose = hand_pose + a["hands_mean"][None]
    full = np.concatenate([np.asarray(rot, np.float64), hand_pose], 1)
    v_shaped = a["v_template"][None] + np.einsum("bl,mkl->bmk", np.asarray(betas, np.float64), a["shapedirs"])
    J = np.einsum("bik,ji->bjk", v_shaped, a["J_regressor"])
    R = rodrigues_np(full.reshape(-1, 3)).reshape(B, 16, 3, 3)
    pf = (R[:, 1:] - np.eye(3)).reshape(B, -1)
    v_posed = v_shaped + (pf @ a["posedirs"].reshape(135, -1)).reshape(B, -1, 3)
    G = np.zeros((B, 16, 4, 4))
    for j in range(16):
        Tm = np.zeros((B, 4, 4))
        Tm[:, :3, :3] = R[:, j]
        Tm[:, :3, 3] = J[:, j] - (J[:, parents[j]] if j > 0 else 0)
        Tm[:, 3, 3] = 1
        G[:, j] = Tm if j == 0 else G[:, parents[j]] @ Tm
    posed_joints = G[:, :, :3, 3].copy()
    A = G.copy()
    A[:, :, :3, 3] -= np.einsum("bjik,bjk->bji", G[:, :, :3, :3], J)
    T = np.einsum("vj,bjik->bvik", a["weights"], A)
    vh = np.concatenate([v_posed, np.ones((B, v_posed.shape[1], 1))], 2)
    verts = np.einsum("bvik,bvk->bvi", T, vh)[:, :, :3]
    return verts, posed_joints


# ----------------------------------------------------------------------------- clip / problem generator
def _axis_angle_to_mat(r):
    return rodrigues_np(np.asarray(r, dtype=np.float64).reshape(-1, 3))


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _square_roi(pts2d, expand=0.3):
    """Square bbox (x, y, b) around 2-D points, expanded by BBOX_EXPANSION_FACTOR
    (/root/reference/homan/constants.py:33)."""
    lo, hi = pts2d.min(0), pts2d.max(0)
    c = (lo + hi) / 2
    b = float((hi - lo).max() * (1 + expand))
    return float(c[0] - b / 2), float(c[1] - b / 2), b


def _k_roi(f, cx, cy, x, y, b):
    """Crop+resize intrinsics normalised by REND_SIZE (/root/reference/homan/pose_optimization.py:247-249,321)."""
    return np.array([[f / b, 0, (cx - x) / b], [0, f / b, (cy - y) / b], [0, 0, 1]], dtype=np.float32)


def project_np(verts, K):
    h = verts @ K.T
    return h[..., :2] / h[..., 2:]


CONFIGS = {
    # name: (P inits, T frames, object, with_hand_losses)
    "cfg1": dict(P=1, T=1, obj="cube"),
    "cfg2": dict(P=16, T=10, obj="ellipsoid500"),
    "cfg3": dict(P=16, T=30, obj="ellipsoid500"),
    "cfg5": dict(P=32, T=30, obj="ellipsoid20k"),
}


def make_object(kind):
    if kind == "cube":
        return make_cube(0.10)
    if kind == "ellipsoid500":
        return make_ellipsoid(25, 10, (0.04, 0.06, 0.04))
    if kind == "ellipsoid20k":
        return make_ellipsoid(100, 100, (0.04, 0.06, 0.04))
    if kind == "ellipsoid80":
        return make_ellipsoid(8, 5, (0.04, 0.06, 0.04))
    if kind.startswith("ellipsoid500@") or kind.startswith("ellipsoid80@"):
        # a clip-specific object: same topology, semi-axes drawn from the number after the @ (cfg4: every clip
        # fits its own object, as the reference does per sample)
        base, seed = kind.split("@")
        rng = np.random.default_rng(int(seed))
        axes = tuple(float(x) for x in rng.uniform(0.03, 0.07, size=3))
        return make_ellipsoid(25, 10, axes) if base == "ellipsoid500" else make_ellipsoid(8, 5, axes)
    raise ValueError(kind)


def make_clip(T, obj="ellipsoid500", seed=0, mano_asset=None, render_fn=None, image_size=640, focal=600.0,
              rend_size=256, occlusion_band=24, side="right"):
    """GT trajectory + targets of one synthetic clip. `render_fn(verts[T,V,3], faces[F,3], K[T,3,3]) ->
    masks [T,rend_size,rend_size] in {0,1}` supplies the target silhouettes (oracle on CPU, product on GPU)."""
    rng = np.random.default_rng(seed)
    asset = mano_asset if mano_asset is not None else make_mano_asset(0, side)
    ov, of = make_object(obj)
    cx = cy = image_size / 2.0
    K_pix = np.array([[focal, 0, cx], [0, focal, cy], [0, 0, 1]], dtype=np.float64)
    # smooth random walks
    def walk(start, step, n):
        out = [np.asarray(start, dtype=np.float64)]
        vel = rng.normal(size=np.shape(start)) * step
        for _ in range(n - 1):
            vel = 0.8 * vel + 0.2 * rng.normal(size=np.shape(start)) * step
            out.append(out[-1] + vel)
        return np.stack(out)

    obj_t = walk([rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), rng.uniform(0.45, 0.65)], 0.004, T)
    R0 = _random_rotation(rng)
    obj_aa = walk(np.zeros(3), 0.03, T)
    obj_R = np.stack([R0 @ _axis_angle_to_mat(a)[0] for a in obj_aa])  # row-vector convention: v @ R
    verts_obj = np.einsum("vk,tkj->tvj", ov.astype(np.float64), obj_R) + obj_t[:, None]
    # hand: MANO-local pose, then rigid placement next to the object (grasp-ish, slight interpenetration)
    pca = walk(rng.normal(size=16) * 0.5, 0.03, T)
    mano_rot = walk(rng.normal(size=3) * 0.3, 0.01, T)
    betas = np.zeros((T, 10))
    mano_trans = np.zeros((T, 3))
    hv_local, _ = mano_forward_np(asset, pca, mano_rot, betas, side=side)
    Rh0 = _random_rotation(rng) if T > 1 else np.eye(3)
    hand_aa = walk(np.zeros(3), 0.02, T)
    hand_R = np.stack([Rh0 @ _axis_angle_to_mat(a)[0] for a in hand_aa])
    hv_rot = np.einsum("tvk,tkj->tvj", hv_local + mano_trans[:, None], hand_R)
    # place the hand centroid ~5.5 cm from the object centre in the image plane
    offs = walk(np.array([-0.055, 0.01, 0.0]), 0.002, T)
    hand_t = obj_t + offs - hv_rot.mean(1)
    verts_hand = hv_rot + hand_t[:, None]
    # ROIs and intrinsics
    camintr_nc = K_pix.copy()
    camintr_nc[:2] /= image_size
    camintr = np.repeat(camintr_nc[None].astype(np.float32), T, 0)
    K_roi_obj, K_roi_hand = [], []
    for t in range(T):
        x, y, b = _square_roi(project_np(verts_obj[t], K_pix))
        K_roi_obj.append(_k_roi(focal, cx, cy, x, y, b))
        x, y, b = _square_roi(project_np(verts_hand[t], K_pix))
        K_roi_hand.append(_k_roi(focal, cx, cy, x, y, b))
    K_roi_obj, K_roi_hand = np.stack(K_roi_obj), np.stack(K_roi_hand)
    clip = dict(T=T, image_size=image_size, rend_size=rend_size, side=side, asset=asset,
                obj_verts_can=ov, obj_faces=of, camintr=camintr, K_roi_obj=K_roi_obj, K_roi_hand=K_roi_hand,
                gt=dict(obj_R=obj_R.astype(np.float32), obj_t=obj_t.astype(np.float32),
                        hand_R=hand_R.astype(np.float32), hand_t=hand_t.astype(np.float32),
                        pca=pca.astype(np.float32), mano_rot=mano_rot.astype(np.float32),
                        betas=betas.astype(np.float32), mano_trans=mano_trans.astype(np.float32),
                        verts_obj=verts_obj.astype(np.float32), verts_hand=verts_hand.astype(np.float32),
                        hand_verts_local=hv_local.astype(np.float32)))
    verts2d = project_np(verts_hand, K_pix) + rng.normal(size=(T, N_HAND_VERTS, 2))
    clip["verts2d"] = verts2d.astype(np.float32)
    if render_fn is not None:
        m_obj = np.asarray(render_fn(clip["gt"]["verts_obj"], of, K_roi_obj), dtype=np.float32)
        m_hand = np.asarray(render_fn(clip["gt"]["verts_hand"], asset["f"], K_roi_hand), dtype=np.float32)
        m_obj = (m_obj > 0.5).astype(np.float32)
        m_hand = (m_hand > 0.5).astype(np.float32)
        if occlusion_band:
            for t in range(T):
                c0 = int(rng.integers(0, rend_size - occlusion_band))
                m_obj[t, :, c0:c0 + occlusion_band] = -1
                c0 = int(rng.integers(0, rend_size - occlusion_band))
                m_hand[t, :, c0:c0 + occlusion_band] = -1
        clip["target_masks_object"] = m_obj
        clip["target_masks_hand"] = m_hand
    return clip


def make_inits(clip, P, seed=0, rot_sigma=0.2, trans_sigma=0.02, pca_sigma=0.5, hand_rot_sigma=0.1,
               hand_trans_sigma=0.01):
    """P perturbed initialisations of one clip: dict of [P,T,...] float32 arrays."""
    rng = np.random.default_rng(seed + 7919)
    gt, T = clip["gt"], clip["T"]
    out = {k: [] for k in ("obj_R", "obj_t", "hand_R", "hand_t", "pca")}
    for _ in range(P):
        dR = _axis_angle_to_mat(rng.normal(size=3) * rot_sigma)[0]
        out["obj_R"].append(np.einsum("tij,jk->tik", gt["obj_R"].astype(np.float64), dR))
        out["obj_t"].append(gt["obj_t"] + rng.normal(size=3) * trans_sigma)
        dRh = _axis_angle_to_mat(rng.normal(size=3) * hand_rot_sigma)[0]
        out["hand_R"].append(np.einsum("tij,jk->tik", gt["hand_R"].astype(np.float64), dRh))
        out["hand_t"].append(gt["hand_t"] + rng.normal(size=3) * hand_trans_sigma)
        out["pca"].append(gt["pca"] + rng.normal(size=16) * pca_sigma)
    out = {k: np.stack(v).astype(np.float32) for k, v in out.items()}
    out["mano_rot"] = np.repeat(gt["mano_rot"][None], P, 0)
    out["mano_trans"] = np.repeat(gt["mano_trans"][None], P, 0)
    out["betas"] = np.repeat(gt["betas"][None], P, 0)
    return out


def reference_inputs(clip, inits, p=0):
    """Problem p in the argument schema of the reference's `optimize_hand_object`
    (lists of per-frame dicts of torch tensors)."""
    import torch

    T = clip["T"]
    t32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))  # noqa: E731
    person, obj = [], []
    faces_hand = torch.from_numpy(clip["asset"]["f"].astype(np.int32))[None]
    tm_o = clip.get("target_masks_object")
    tm_h = clip.get("target_masks_hand")
    R = clip["rend_size"]
    for t in range(T):
        person.append({
            "translations": t32(inits["hand_t"][p, t]).view(1, 1, 3),
            "rotations": t32(inits["hand_R"][p, t]).view(1, 3, 3),
            "hand_side": [clip["side"]],
            "faces": faces_hand,
            "mano_trans": t32(inits["mano_trans"][p, t]).view(1, 3),
            "mano_rot": t32(inits["mano_rot"][p, t]).view(1, 3),
            "mano_betas": t32(inits["betas"][p, t]).view(1, 10),
            "mano_pca_pose": t32(inits["pca"][p, t]).view(1, -1),
            "target_masks": t32(tm_h[t]).view(1, R, R) if tm_h is not None else torch.zeros(1, R, R),
            "masks": torch.zeros(1, 8, 8, dtype=torch.bool),
            "verts": t32(clip["gt"]["hand_verts_local"][t]).view(1, N_HAND_VERTS, 3),
            "verts2d": t32(clip["verts2d"][t]).view(1, N_HAND_VERTS, 2),
            "K_roi": t32(clip["K_roi_hand"][t]).view(1, 3, 3),
            "cams": torch.ones(1, 3),
        })
        obj.append({
            "translations": t32(inits["obj_t"][p, t]).view(1, 1, 3),
            "rotations": t32(inits["obj_R"][p, t]).view(1, 3, 3),
            "target_masks": t32(tm_o[t]).view(1, R, R) if tm_o is not None else torch.zeros(1, R, R),
            "full_mask": torch.zeros(8, 8, dtype=torch.bool),
            "K_roi": t32(clip["K_roi_obj"][t]).view(1, 1, 3, 3),
        })
    return {
        "person_parameters": person,
        "object_parameters": obj,
        "objvertices": np.repeat(clip["obj_verts_can"][None], T, 0),
        "objfaces": np.repeat(clip["obj_faces"][None], T, 0),
        "camintr": clip["camintr"],
        "image_size": clip["image_size"],
    }


def default_loss_weights(**overrides):
    """Every `lw_*` key HOMan.forward reads (/root/reference/homan/homan.py:433-506), all zero."""
    lw = {k: 0.0 for k in ("lw_pca", "lw_smooth_hand", "lw_smooth_obj", "lw_collision", "lw_contact",
                           "lw_v2d_hand", "lw_sil_obj", "lw_sil_hand", "lw_inter", "lw_scale_obj",
                           "lw_scale_hand", "lw_depth")}
    lw.update(overrides)
    return lw


def step1_loss_weights():
    """Defaults of /root/reference/fit_vid_dataset.py:91-158 (step 1)."""
    return default_loss_weights(lw_pca=0.004, lw_smooth_hand=2000.0, lw_smooth_obj=2000.0, lw_v2d_hand=50.0,
                                lw_sil_obj=1.0, lw_inter=1.0, lw_scale_obj=0.001, lw_scale_hand=0.001)


def step2_loss_weights():
    """Step 2 (README.md:217): step 1 + collision 0.001 + contact 1."""
    return default_loss_weights(lw_pca=0.004, lw_smooth_hand=2000.0, lw_smooth_obj=2000.0, lw_v2d_hand=50.0,
                                lw_sil_obj=1.0, lw_inter=1.0, lw_scale_obj=0.001, lw_scale_hand=0.001,
                                lw_collision=0.001, lw_contact=1.0)


def make_batch(clip, inits):
    """Problem-major batch (P problems x T frames) of one clip and its P initialisations.
    Every per-frame array is [P, T, ...]; targets are replicated per problem."""
    P = inits["obj_t"].shape[0]
    rep = lambda x: np.ascontiguousarray(np.repeat(np.asarray(x)[None], P, 0))  # noqa: E731
    batch = {
        "P": P, "T": clip["T"], "side": clip["side"], "image_size": clip["image_size"],
        "mano_asset": clip["asset"],
        "obj_verts_can": clip["obj_verts_can"], "obj_faces": clip["obj_faces"], "hand_faces": clip["asset"]["f"],
        "camintr": rep(clip["camintr"]), "K_roi_obj": rep(clip["K_roi_obj"]), "K_roi_hand": rep(clip["K_roi_hand"]),
        "verts2d": rep(clip["verts2d"]),
        "target_masks_object": rep(clip["target_masks_object"]),
        "target_masks_hand": rep(clip["target_masks_hand"]),
    }
    for k in ("obj_R", "obj_t", "hand_R", "hand_t", "pca", "mano_rot", "mano_trans", "betas"):
        batch[k] = np.ascontiguousarray(inits[k])
    return batch


def concat_batches(batches):
    """Stacks the problems of several clips (same T, Vo, Fo and MANO asset; cfg4) into one clip-major batch: every
    clip keeps its own object mesh (obj_verts_can [C,Vo,3], obj_faces [C,Fo,3], clip_of_problem [P])."""
    out = dict(batches[0])
    for k, v in batches[0].items():
        if isinstance(v, np.ndarray) and k not in ("obj_verts_can", "obj_faces", "hand_faces"):
            out[k] = np.concatenate([b[k] for b in batches], 0)
    out["obj_verts_can"] = np.stack([np.asarray(b["obj_verts_can"]) for b in batches])
    out["obj_faces"] = np.stack([np.asarray(b["obj_faces"]) for b in batches])
    out["clip_of_problem"] = np.concatenate([np.full(b["P"], c, np.int64) for c, b in enumerate(batches)])
    out["P"] = sum(b["P"] for b in batches)
    return out


