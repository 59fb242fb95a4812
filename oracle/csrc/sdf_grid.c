/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement of the `sdf` CUDA extension of hassony2/multiperson that the
 * reference calls at /root/reference/homan/interactions/scenesdf.py:32,119
 * (`SDF()(faces, vertices)` -> phi [B,G,G,G]). The package is not vendored in the
 * reference (README.md:59-60); semantics follow SURVEY.md Appendix A.4:
 * "parity unpinned".
 *
 *   voxel (k, j, i) = (z, y, x), x fastest; centre c = -1 + (idx + 0.5) * 2 / G
 *   phi = +min_f dist(c, triangle f) when an axis ray from c crosses the mesh an
 *         odd number of times (inside), else -min_f dist (outside).
 * The sign predicate is evaluated without FMA contraction (-ffp-contract=off);
 * the CUDA kernel uses __fmul_rn/__fadd_rn for the same expressions.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* squared distance from p to triangle (a, b, c): closest-point region walk */
static float point_tri_dist2(const float *p, const float *a, const float *b, const float *c) {
    float ab[3], ac[3], ap[3], bp[3], cp[3], q[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return dot3(ap, ap);
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return dot3(bp, bp);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
        for (int k = 0; k < 3; ++k) q[k] = ap[k] - v * ab[k];
        return dot3(q, q);
    }
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return dot3(cp, cp);
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float w = d2 / (d2 - d6);
        for (int k = 0; k < 3; ++k) q[k] = ap[k] - w * ac[k];
        return dot3(q, q);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; ++k) q[k] = bp[k] - w * (c[k] - b[k]);
        return dot3(q, q);
    }
    const float denom = 1.f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    for (int k = 0; k < 3; ++k) q[k] = ap[k] - (ab[k] * v + ac[k] * w);
    return dot3(q, q);
}

/* does the ray p + s * (+x), s > 0 cross triangle (a, b, c)?  (projection on the yz plane) */
static inline int ray_x_hits(const float *p, const float *a, const float *b, const float *c) {
    const float w0 = (c[1] - b[1]) * (p[2] - b[2]) - (c[2] - b[2]) * (p[1] - b[1]);
    const float w1 = (a[1] - c[1]) * (p[2] - c[2]) - (a[2] - c[2]) * (p[1] - c[1]);
    const float w2 = (b[1] - a[1]) * (p[2] - a[2]) - (b[2] - a[2]) * (p[1] - a[1]);
    const float area = w0 + w1 + w2;
    int inside;
    if (area > 0.f) inside = (w0 >= 0.f && w1 >= 0.f && w2 >= 0.f);
    else if (area < 0.f) inside = (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
    else return 0;
    if (!inside) return 0;
    const float xs = (w0 * a[0] + w1 * b[0] + w2 * c[0]) / area;
    return xs > p[0];
}

/* faces [F,3] int32 (shared by the batch), verts [B,V,3] in [-1,1]^3, phi [B,G,G,G] */
void sdf_grid(const int32_t *faces, int F, const float *verts, int B, int V, int G, float *phi) {
    #pragma omp parallel for schedule(dynamic, 64) collapse(2)
    for (int b = 0; b < B; ++b)
        for (int vox = 0; vox < G * G * G; ++vox) {
            const int i = vox % G, j = (vox / G) % G, k = vox / (G * G);
            const float *vb = verts + (size_t)b * V * 3;
            float c[3];
            c[0] = -1.f + (i + 0.5f) * 2.f / G;
            c[1] = -1.f + (j + 0.5f) * 2.f / G;
            c[2] = -1.f + (k + 0.5f) * 2.f / G;
            float best = INFINITY;
            int hits = 0;
            for (int f = 0; f < F; ++f) {
                const float *va = vb + 3 * faces[3 * f], *vbb = vb + 3 * faces[3 * f + 1], *vc = vb + 3 * faces[3 * f + 2];
                const float d2 = point_tri_dist2(c, va, vbb, vc);
                if (d2 < best) best = d2;
                hits += ray_x_hits(c, va, vbb, vc);
            }
            const float d = sqrtf(best);
            phi[(size_t)b * G * G * G + vox] = (hits & 1) ? d : -d;
        }
}
