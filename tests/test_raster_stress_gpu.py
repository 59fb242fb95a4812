"""GPU: adversarial inputs for the silhouette rasteriser against the CPU oracle (oracle/nmr.py + nmr_raster.c).

The forward kernel does not test every sample with the reference predicate and does not evaluate the reference depth
for every sample: it derives exact per-row coverage intervals from the monotonicity of the edge predicate and ranks
samples with a fast depth, re-resolving near-ties with the reference arithmetic. These cases aim at the corners of
that construction: coincident and interpenetrating layers (depth ties), vertices on pixel centres and on integer
pixel coordinates (edge ties), degenerate and sub-pixel faces, geometry crossing the near plane, behind the camera and
off screen, a non-power-of-two raster, per-image face lists. Bar: face_index and alpha bit-exact, gradient 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle(ndc, faces_b, image_size, aa, grad_alpha=None):
    from oracle import nmr
    ndc = ndc.clone().requires_grad_(grad_alpha is not None)
    f2 = torch.cat((faces_b, faces_b[:, :, [2, 1, 0]]), dim=1)
    fv = nmr.vertices_to_faces(ndc, f2)
    alpha, fi = nmr.rasterize_silhouettes(fv, image_size, aa, return_face_index=True)
    g = None
    if grad_alpha is not None:
        alpha.backward(grad_alpha)
        g = ndc.grad
    return alpha.detach(), fi, g


def _check(ndc, faces, image_size, aa, with_grad=True, faces_batched=False):
    from homan_b200 import ops
    ndc = torch.as_tensor(ndc, dtype=torch.float32).contiguous()
    faces = torch.as_tensor(faces)
    fb = faces.long() if faces_batched else faces.long()[None].repeat(ndc.shape[0], 1, 1)
    a_ref, fi_ref, _ = _oracle(ndc, fb, image_size, aa)
    nd = ndc.cuda().requires_grad_()
    fd = (faces if faces_batched else faces[None]).int().cuda().contiguous()
    alpha, fi = ops.rasterize_silhouettes(nd, fd, image_size, aa, return_face_index=True)
    bad = int((fi.cpu() != fi_ref).sum())
    assert bad == 0, f"face_index differs at {bad} pixels"
    assert torch.equal(alpha.cpu(), a_ref)
    if not with_grad:
        return
    target = torch.roll(a_ref, shifts=(5, -7), dims=(1, 2)).round()
    g_alpha = 2 * (a_ref - target) / target[0].numel()
    if float(g_alpha.abs().max()) == 0:
        return
    _, _, g_ref = _oracle(ndc, fb, image_size, aa, g_alpha)
    alpha.backward(g_alpha.cuda())
    scale = float(g_ref.abs().max())
    if scale > 0:
        assert float((nd.grad.cpu() - g_ref).abs().max()) <= 1e-4 * scale


def _soup(rng, n_faces, size, z_lo=0.3, z_hi=2.0):
    """Random triangles: independent vertices, extent `size` in NDC."""
    c = rng.uniform(-0.9, 0.9, size=(n_faces, 1, 2))
    xy = c + rng.uniform(-size, size, size=(n_faces, 3, 2))
    z = rng.uniform(z_lo, z_hi, size=(n_faces, 3, 1))
    verts = np.concatenate((xy, z), 2).reshape(-1, 3).astype(np.float32)
    faces = np.arange(3 * n_faces).reshape(n_faces, 3)
    return verts, faces


@pytest.mark.parametrize("aa", [True, False])
def test_random_triangle_soup_with_coincident_layers(aa):
    rng = np.random.default_rng(1)
    v, f = _soup(rng, 300, 0.25)
    # duplicate a third of the faces exactly (depth ties -> lowest face index) and a third at a depth a few ulp off
    dup = v.reshape(-1, 3, 3)[:100].copy()
    near = v.reshape(-1, 3, 3)[100:200].copy()
    near[:, :, 2] *= np.float32(1 + 3e-7)
    verts = np.concatenate((v.reshape(-1, 3, 3), dup, near)).reshape(-1, 3)
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    ndc = np.stack((verts, verts * np.array([1, -1, 1], np.float32)))
    _check(ndc, faces, 128, aa)


def test_vertices_on_pixel_centres_and_integer_coordinates():
    """is = 256 (aa off): pixel centres sit at (2i + 1 - is) / is, pixel coordinate i; integer pixel coordinates also
    exercise the p0.d0 == d0 branches of backward_pixel_map."""
    rng = np.random.default_rng(2)
    S = 256
    n = 200
    pix = rng.integers(8, S - 8, size=(n, 3, 2)).astype(np.float64)
    pix[::2] += 0.0           # vertices exactly on pixel centres (pixel coordinate integer)
    pix[1::2] += 0.5          # vertices exactly on pixel corners
    xy = (2 * pix + 1 - S) / S
    z = rng.uniform(0.5, 1.5, size=(n, 3, 1))
    verts = np.concatenate((xy, z), 2).reshape(-1, 3).astype(np.float32)
    faces = np.arange(3 * n).reshape(n, 3)
    _check(verts[None], faces, S, False)


def test_degenerate_subpixel_and_sliver_faces():
    rng = np.random.default_rng(3)
    v, f = _soup(rng, 200, 0.004)                      # sub-pixel triangles
    v2, f2 = _soup(rng, 100, 0.3)
    v2 = v2.reshape(-1, 3, 3)
    v2[:, 2, :2] = v2[:, 0, :2] + (v2[:, 1, :2] - v2[:, 0, :2]) * 0.5 + rng.normal(size=(100, 2)) * 1e-4  # slivers
    v3 = v2[:20].copy()
    v3[:, 1] = v3[:, 0]                                # two identical vertices: zero area
    verts = np.concatenate((v.reshape(-1, 3, 3), v2, v3)).reshape(-1, 3).astype(np.float32)
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    _check(verts[None], faces, 128, True)


def test_faces_whose_two_windings_are_both_front_facing():
    """Signed area exactly 0 in fp32: neither the face nor its fill_back copy is back-face culled, the reference
    rasterises and differentiates BOTH (f and F + f). Zero-area faces (repeated vertex, exactly collinear vertices,
    vertices on one row of pixel centres) and slivers whose fp32 cross product cancels although the area does not."""
    rng = np.random.default_rng(6)
    S = 256
    tris = []
    for _ in range(30):   # repeated vertex
        a, b = rng.uniform(-0.8, 0.8, 2), rng.uniform(-0.8, 0.8, 2)
        tris.append([a, a, b])
    for _ in range(30):   # exactly collinear, dyadic coordinates
        a = np.round(rng.uniform(-0.7, 0.7, 2) * 64) / 64
        d = np.round(rng.uniform(-0.1, 0.1, 2) * 256) / 256
        tris.append([a, a + d, a + 2 * d])
    for _ in range(20):   # on one row of pixel centres
        yc = (2 * rng.integers(20, S - 20) + 1 - S) / S
        xs = np.sort(rng.uniform(-0.8, 0.8, 3))
        tris.append([[xs[0], yc], [xs[1], yc], [xs[2], yc]])
    found = 0
    while found < 40:     # slivers with a cancelling fp32 cross product
        a, b = rng.uniform(-0.8, 0.8, 2).astype(np.float32), rng.uniform(-0.8, 0.8, 2).astype(np.float32)
        t = np.float32(rng.uniform(0.2, 0.8))
        c = (a + (b - a) * t).astype(np.float32)
        c[1] = np.nextafter(c[1], np.float32(2.0)) if found % 2 else c[1]
        lhs = np.float32(c[1] - a[1]) * np.float32(b[0] - a[0])
        rhs = np.float32(b[1] - a[1]) * np.float32(c[0] - a[0])
        lhs2 = np.float32(a[1] - c[1]) * np.float32(b[0] - c[0])
        rhs2 = np.float32(b[1] - c[1]) * np.float32(a[0] - c[0])
        if not (lhs < rhs) and not (lhs2 < rhs2) and (c != a).any():
            tris.append([a, b, c])
            found += 1
    xy = np.asarray([[np.asarray(p, dtype=np.float64) for p in t] for t in tris], dtype=np.float32)
    z = rng.uniform(0.5, 1.5, size=(xy.shape[0], 3, 1)).astype(np.float32)
    big, _ = _soup(rng, 60, 0.3)      # ordinary faces around them so that the loss gradient is not empty
    verts = np.concatenate((np.concatenate((xy, z), 2).reshape(-1, 3), big)).astype(np.float32)
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    _check(verts[None], faces, S // 2, True)
    _check(verts[None], faces, S, False)


def test_near_plane_behind_camera_and_off_screen():
    rng = np.random.default_rng(4)
    v, f = _soup(rng, 150, 0.4, z_lo=-0.5, z_hi=0.6)   # straddles z = 0 and the near plane (0.1)
    v[::7, 2] = 0.1                                    # vertices exactly on the near plane
    v[::11, :2] *= 3.0                                 # far off screen
    v2, _ = _soup(rng, 20, 0.2, z_lo=99.9, z_hi=100.1)  # around the far plane
    verts = np.concatenate((v, v2))
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    _check(verts[None], faces, 128, True, with_grad=False)
    _check(verts[None], faces, 128, False, with_grad=False)


def test_non_power_of_two_raster_and_per_image_faces():
    rng = np.random.default_rng(5)
    v, f = _soup(rng, 120, 0.3)
    ndc = np.stack((v, v[::-1].copy(), v * np.array([-1, 1, 1], np.float32)))
    faces_b = np.stack((f, f[::-1].copy(), np.roll(f, 1, axis=1)))
    _check(ndc, faces_b, 96, True, faces_batched=True)   # raster 192 = 3 tiles per side, pixel centres not dyadic
    _check(ndc, faces_b, 192, False, faces_batched=True)


def test_interpenetrating_closed_meshes():
    from homan_b200 import synth
    va, fa = synth.make_object("ellipsoid500")
    vb = va[:, [1, 0, 2]] * np.array([1.0, 1.0, 1.0]) + np.array([0.01, 0.0, 0.0])
    verts = np.concatenate((va, vb)).astype(np.float32) * 8 + np.array([0, 0, 1.2], np.float32)
    faces = np.concatenate((fa, fa[:, ::-1] + va.shape[0]))
    ndc = verts.copy()
    ndc[:, :2] = verts[:, :2] / verts[:, 2:3] * 1.5
    _check(ndc[None], faces, 256, True)


def test_empty_inputs():
    from homan_b200 import ops
    ndc = torch.zeros(0, 5, 3, device="cuda")
    faces = torch.zeros(1, 2, 3, dtype=torch.int32, device="cuda")
    alpha = ops.rasterize_silhouettes(ndc, faces, 64, True)
    assert alpha.shape == (0, 64, 64)
    ndc = torch.rand(2, 5, 3, device="cuda")
    faces = torch.zeros(1, 0, 3, dtype=torch.int32, device="cuda")
    alpha, fi = ops.rasterize_silhouettes(ndc, faces, 64, True, return_face_index=True)
    assert float(alpha.abs().max()) == 0 and int(fi.max()) == -1


def test_closed_mesh_across_the_image_border():
    """A closed mesh straddling the left / bottom image border: faces just outside the image (a corner coordinate in
    (-1, 0) pixels) still get scan-line 0 from the reference's truncating range arithmetic, with sweep ends
    extrapolated outside the triangle - their in-sweeps contribute although their pixel box is fully covered."""
    from homan_b200 import synth
    ov, of = synth.make_object("ellipsoid500")
    rng = np.random.default_rng(12)
    ndcs = []
    for k in range(6):
        s = rng.uniform(6.0, 9.0)
        off = np.array([-1.0 + rng.uniform(-0.15, 0.15), -1.0 + rng.uniform(-0.15, 0.15) if k % 2 else rng.uniform(-0.5, 0.5)])
        xy = ov[:, :2] * s + off
        ndcs.append(np.concatenate((xy, ov[:, 2:3] * s + 2.0), 1))
    ndc = np.stack(ndcs).astype(np.float32)
    px = 0.5 * (ndc[..., :2] * 256 + 255)
    assert ((px > -1) & (px < 0)).any()   # the case is present
    _check(ndc, of, 128, True)
    _check(ndc[:2], of, 256, False)


@pytest.mark.parametrize("aa", [True, False])
def test_needles_and_the_tight_forward_box(aa):
    """The forward visits only the rows / columns of floor(min) .. ceil(max) of a face's corners unless the face is a
    needle (|det| < 1e-3 |longest edge|^2, raster.cu face_setup_kernel), for which it keeps the reference's box with a
    pixel of slack. Needles of every aspect ratio around that threshold, at random orientations, with tips and edges
    landing between, on and next to pixel centres, over a backdrop so that depth ranking runs too."""
    rng = np.random.default_rng(17)
    tris, n = [], 0
    for aspect in (3.0, 30.0, 300.0, 1e3, 3e3, 1e4, 1e5, 1e6, 1e7):
        for _ in range(40):
            c = rng.uniform(-0.8, 0.8, size=2)
            th = rng.uniform(0, 2 * np.pi)
            L = rng.uniform(0.05, 0.9)
            d = np.array([np.cos(th), np.sin(th)])
            nrm = np.array([-d[1], d[0]])
            a = c - 0.5 * L * d
            b = c + 0.5 * L * d + rng.uniform(-1, 1) * (L / aspect) * nrm
            tip = c + rng.uniform(-0.5, 0.5) * L * d + (L / aspect) * nrm
            z = rng.uniform(0.5, 1.5, size=(3, 1))
            tris.append(np.concatenate((np.stack((a, b, tip)), z), 1))
            n += 1
    # snap some corners onto pixel centres / integer pixel coordinates of the raster
    S = 256 if aa else 128
    t = np.stack(tris)
    snap = rng.random(t.shape[:2]) < 0.2
    px = np.round(0.5 * (t[..., :2] * S + S - 1))
    t[..., :2] = np.where(snap[..., None], (2 * px + 1 - S) / S, t[..., :2])
    backdrop = np.array([[[-0.9, -0.9, 2.0], [0.9, -0.9, 2.0], [0.0, 0.9, 2.0]]])
    verts = np.concatenate((t, backdrop)).reshape(-1, 3).astype(np.float32)
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    ndc = np.stack((verts, verts * np.array([-1, 1, 1], np.float32)))
    _check(ndc, faces, 128, aa)


@pytest.mark.parametrize("aa", [True, False])
def test_dense_soup_of_small_and_large_faces(aa):
    """A soup of 5000 triangles of mixed sizes - mostly a few pixels, some spanning tiles -, interleaved in the tile
    lists, with exact duplicates and near-duplicates (depth ties), vertices snapped onto pixel centres, faces crossing
    the near plane and the image border: several batches per tile list, long hidden-layer lists (thread-per-entry
    filter), the lean path of untouched tiles next to crowded ones."""
    rng = np.random.default_rng(23)
    small_v, _ = _soup(rng, 4200, 0.02)
    mid_v, _ = _soup(rng, 600, 0.08)
    big_v, _ = _soup(rng, 40, 0.6)
    t = np.concatenate((small_v, mid_v, big_v)).reshape(-1, 3, 3)
    rng.shuffle(t)   # small and large faces interleaved in the tile lists
    t[:50, :, 2] = rng.uniform(0.05, 0.15, size=(50, 3))          # across the near plane (0.1)
    t[50:120, :, :2] += np.sign(t[50:120, :1, :2]) * 0.35          # across the image border
    S = 256 if aa else 128
    snap = rng.random(t.shape[:2]) < 0.1
    px = np.round(0.5 * (t[..., :2] * S + S - 1))
    t[..., :2] = np.where(snap[..., None], (2 * px + 1 - S) / S, t[..., :2])
    dup = t[200:400].copy()
    near = t[400:600].copy()
    near[:, :, 2] *= np.float32(1 + 3e-7)
    verts = np.concatenate((t, dup, near)).reshape(-1, 3).astype(np.float32)
    faces = np.arange(verts.shape[0]).reshape(-1, 3)
    assert faces.shape[0] >= 4096
    ndc = np.stack((verts, verts * np.array([1, -1, 1], np.float32)))
    _check(ndc, faces, 128, aa)
