"""Generates homan_b200/data/hand_template.npz: the template surface of the synthetic MANO-shaped hand with
well-shaped faces (778 vertices, 1552 watertight faces, of which the first 1538 are the open raster mesh).

MANO itself is licence-gated and absent (/root/reference/README.md:72-90); the benchmark hand only has to have MANO's
size (778 v / 1538 f / 1552 closed f) and a realistic triangulation. Construction: the planar "mitten" outline of
homan_b200/synth.py (palm + five finger lobes) is sampled at spacing ~h, its interior filled with a hexagonal grid of the
same spacing, and the 2-D Delaunay triangulation of both (restricted to the outline) becomes the top and the bottom
sheet of a thin pillow that share the outline vertices: V = 2 N_interior + N_outline, F = 2 F_2D = 2 V - 4.
h is searched so that V = 778 exactly. The 14 faces nearest the wrist go last (the open mesh drops them, as MANO's
open wrist does).

    python scripts/make_hand_template.py
"""
import math
import os
import sys

import numpy as np
from scipy.spatial import Delaunay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from homan_b200 import synth  # noqa: E402

NV, NF_CLOSED, NF_OPEN = 778, 1552, 1538
THICK = 0.011


def outline_points(n_dense=20000):
    phi = np.linspace(0, 2 * math.pi, n_dense, endpoint=False)
    r = synth._hand_outline(phi)
    return np.stack([r * np.cos(phi), r * np.sin(phi)], 1)


def resample_closed(poly, n):
    seg = np.linalg.norm(np.roll(poly, -1, 0) - poly, axis=1)
    s = np.concatenate([[0], np.cumsum(seg)])
    t = np.arange(n) * s[-1] / n
    k = np.searchsorted(s, t, side="right") - 1
    a = (t - s[k]) / seg[k]
    return poly[k] * (1 - a[:, None]) + poly[(k + 1) % len(poly)] * a[:, None], s[-1]


def inside(poly_phi_r, pts):
    phi = np.arctan2(pts[:, 1], pts[:, 0])
    return np.hypot(pts[:, 0], pts[:, 1]) / synth._hand_outline(phi)   # < 1 inside (star-shaped outline)


def build(h):
    dense = outline_points()
    perim = np.linalg.norm(np.roll(dense, -1, 0) - dense, axis=1).sum()
    # interior: hexagonal grid, kept away from the outline
    xs = np.arange(-0.2, 0.2, h)
    ys = np.arange(-0.2, 0.2, h * math.sqrt(3) / 2)
    gx, gy = np.meshgrid(xs, ys)
    gx = gx + (np.arange(len(ys)) % 2)[:, None] * h / 2
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    rel = inside(None, pts)
    pts = pts[rel < 1]
    # distance to the outline
    from scipy.spatial import cKDTree
    d, _ = cKDTree(dense).query(pts)
    interior = pts[d > 0.62 * h]
    n_bnd = NV - 2 * len(interior)
    return interior, n_bnd, perim, dense


def triangulate(bnd, interior):
    p2 = np.concatenate([bnd, interior])
    tri = Delaunay(p2).simplices
    tri = tri[inside(None, p2[tri].mean(1)) < 1.0]
    return p2, tri


def chords(tri, nb):
    """Interior edges that join two outline vertices: the top and the bottom sheet would share them (non-manifold)."""
    e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), 1)
    e = np.unique(e[(e[:, 1] < nb)], axis=0)
    d = (e[:, 1] - e[:, 0]) % nb
    return e[(d != 1) & (d != nb - 1)]


def make(h):
    """-> (outline points, interior points, triangles) with 2 * interior + outline = NV and no chord, or None."""
    interior, n_bnd, perim, dense = build(h)
    for _ in range(20):
        n_bnd = NV - 2 * len(interior)
        if n_bnd < 16:
            return None
        bnd, perim = resample_closed(dense, n_bnd)
        p2, tri = triangulate(bnd, interior)
        ch = chords(tri, n_bnd)
        if len(ch) == 0:
            if len(tri) != 2 * (n_bnd + len(interior)) - 2 - n_bnd:
                return None
            return bnd, interior, tri, perim / n_bnd
        interior = np.concatenate([interior, 0.5 * (p2[ch[:, 0]] + p2[ch[:, 1]])])   # split every chord at its midpoint
    return None


def main():
    best = None
    for h in np.linspace(0.0040, 0.0080, 401):
        m = make(h)
        if m is None:
            continue
        score = abs(math.log(m[3] / h))
        if best is None or score < best[0]:
            best = (score, h) + m
    _, h, bnd, interior, tri, hb = best
    n_bnd = len(bnd)
    print(f"h = {h * 1e3:.3f} mm, interior {len(interior)}, outline {n_bnd} (spacing {hb * 1e3:.3f} mm)")
    p2 = np.concatenate([bnd, interior])
    # counter-clockwise
    a = p2[tri]
    u, w = a[:, 1] - a[:, 0], a[:, 2] - a[:, 0]
    area2 = u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]
    tri[area2 < 0] = tri[area2 < 0][:, ::-1]
    nb, ni = len(bnd), len(interior)
    # pillow: z = +-THICK * (1 - rel^2)^0.5 ; top sheet = interior copy A, bottom sheet = interior copy B
    rel = np.clip(inside(None, interior), 0, 1)
    z = THICK * np.sqrt(1 - rel ** 2)
    verts = np.concatenate([np.concatenate([bnd, np.zeros((nb, 1))], 1),
                            np.concatenate([interior, z[:, None]], 1),
                            np.concatenate([interior, -z[:, None]], 1)])
    top = tri.copy()
    bot = tri[:, ::-1].copy()
    bot[bot >= nb] += ni
    faces = np.concatenate([top, bot]).astype(np.int64)
    assert verts.shape == (NV, 3) and faces.shape == (NF_CLOSED, 3), (verts.shape, faces.shape)
    vol = np.einsum("ij,ij->i", verts[faces[:, 0]], np.cross(verts[faces[:, 1]], verts[faces[:, 2]])).sum()
    assert vol > 0
    edges = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), 1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    assert (cnt == 2).all(), "not watertight"
    # wrist cap: the 14 faces nearest the rim point opposite the fingers go last
    rw = synth._hand_outline(np.array([math.pi]))[0]
    wrist = np.array([-rw, 0.0, 0.0])
    dist = ((verts[faces].mean(1) - wrist) ** 2).sum(1)
    cap = np.argsort(dist, kind="stable")[:NF_CLOSED - NF_OPEN]
    keep = np.ones(NF_CLOSED, bool)
    keep[cap] = False
    faces = np.concatenate([faces[keep], faces[~keep]])
    e = np.stack([np.linalg.norm(verts[faces[:, k]] - verts[faces[:, (k + 1) % 3]], axis=1) for k in range(3)], 1)
    r = e.max(1) / e.min(1)
    print("edge length mean %.4f  min %.4f  max %.4f; longest/shortest edge of a face: worst %.2f, median %.2f, p90 %.2f" %
          (e.mean(), e.min(), e.max(), r.max(), np.median(r), np.percentile(r, 90)))
    out = os.path.join(ROOT, "homan_b200", "data", "hand_template.npz")
    np.savez_compressed(out, verts=verts.astype(np.float32), closed_faces=faces.astype(np.int16))
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
