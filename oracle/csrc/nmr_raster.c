/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU restatement, in plain C, of the silhouette rasteriser of `neural_renderer`
 * (Kato-style NMR, PyTorch port used by hassony2/multiperson), which the reference
 * calls at /root/reference/homan/losses.py:73-77,187 and homan/homan.py:168-176.
 * The package is NOT vendored in the reference (README.md:54-58, unpinned git
 * clone), so this file restates its published algorithm as recorded in
 * "parity unpinned" for the third-party semantics.
 *
 *   nmr_face_index_map : forward_face_index_map (pixel-parallel, nearest
 *                        front-facing face, ties -> lowest face index)
 *   nmr_alpha_flip_pool: alpha = face_index >= 0, vertical flip, 2x2 avg-pool
 *   nmr_pixel_map_bwd  : backward_pixel_map (approximate d loss / d face xy)
 *
 * All arithmetic is fp32 with no FMA contraction (build with -ffp-contract=off)
 * so that the CUDA kernels, which use __fmul_rn/__fadd_rn for the same
 * predicates, can reproduce coverage bit for bit.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* float -> int conversion with CUDA cvt.rzi.s32.f32 semantics (NaN -> 0, saturating). */
static inline int f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)v;
}

/* Named knob (parity unpinned, like the raster eps): upstream's face set-up reportedly keeps the barycentric
 * determinant away from zero (|det| >= 1e-10) before dividing by it. Only faces that are degenerate to that degree
 * are affected (their samples fail the near / far test either way: inf or NaN depth). 0 = divide by the raw
 * determinant. The CUDA kernel carries the same constant (HM_RASTER_DET_CLAMP in homan_b200/csrc/raster.cu). */
#ifndef NMR_DET_CLAMP
#define NMR_DET_CLAMP 0.f
#endif
static inline float clamp_det(float den) {
    if (NMR_DET_CLAMP > 0.f) {
        if (den > 0.f) return den < NMR_DET_CLAMP ? NMR_DET_CLAMP : den;
        return den > -NMR_DET_CLAMP ? -NMR_DET_CLAMP : den;
    }
    return den;
}

static inline int is_backface(const float *f) {
    /* f = x0 y0 z0 x1 y1 z1 x2 y2 z2 */
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

/* faces [B, nf, 9] (already doubled when fill_back), outputs [B, is, is]. */
void nmr_face_index_map(const float *faces, int B, int nf, int is, float near_, float far_,
                        int32_t *face_index, float *depth_map) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *fb = faces + (size_t)b * nf * 9;
        int32_t *fi = face_index + (size_t)b * is * is;
        float *dm = depth_map ? depth_map + (size_t)b * is * is : 0;
        for (int i = 0; i < is * is; ++i) { fi[i] = -1; if (dm) dm[i] = far_; }
        /* face-major loop over the face's pixel bounding box: visits exactly the
         * (pixel, face) pairs that can pass the inside test; the winner rule
         * (smaller depth, ties -> smaller face index) makes it order independent
         * and equal to the upstream pixel-major loop. */
        float *zb = dm;
        float *ztmp = 0;
        if (!zb) { /* private depth buffer */
            static _Thread_local float *buf = 0; static _Thread_local int cap = 0;
            if (cap < is * is) { buf = (float *)__builtin_realloc(buf, sizeof(float) * is * is); cap = is * is; }
            ztmp = buf; zb = ztmp;
            for (int i = 0; i < is * is; ++i) zb[i] = far_;
        }
        for (int fn = 0; fn < nf; ++fn) {
            const float *f = fb + (size_t)fn * 9;
            if (is_backface(f)) continue;
            float p[3][2];
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 2; ++d) p[k][d] = 0.5f * (f[3 * k + d] * is + is - 1);
            float inv[9] = {
                p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
            float den = p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) +
                        p[1][0] * (p[2][1] - p[0][1]);
            den = clamp_det(den);
            for (int k = 0; k < 9; ++k) inv[k] /= den;
            /* conservative pixel bbox (1 px slack each side) */
            float xmin = fminf(fminf(p[0][0], p[1][0]), p[2][0]), xmax = fmaxf(fmaxf(p[0][0], p[1][0]), p[2][0]);
            float ymin = fminf(fminf(p[0][1], p[1][1]), p[2][1]), ymax = fmaxf(fmaxf(p[0][1], p[1][1]), p[2][1]);
            if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) continue;
            int x0 = f2i(floorf(xmin)) - 1, x1 = f2i(ceilf(xmax)) + 1;
            int y0 = f2i(floorf(ymin)) - 1, y1 = f2i(ceilf(ymax)) + 1;
            if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0;
            if (x1 > is - 1) x1 = is - 1; if (y1 > is - 1) y1 = is - 1;
            for (int yi = y0; yi <= y1; ++yi) {
                const float yp = (float)(2 * yi + 1 - is) / is;
                for (int xi = x0; xi <= x1; ++xi) {
                    const float xp = (float)(2 * xi + 1 - is) / is;
                    if ((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) continue;
                    if ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) continue;
                    if ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])) continue;
                    float w[3], ws = 0.f;
                    for (int k = 0; k < 3; ++k) {
                        w[k] = inv[3 * k] * xi + inv[3 * k + 1] * yi + inv[3 * k + 2];
                        w[k] = fminf(fmaxf(w[k], 0.f), 1.f);
                        ws += w[k];
                    }
                    for (int k = 0; k < 3; ++k) w[k] /= ws;
                    const float zp = 1.f / (w[0] / f[2] + w[1] / f[5] + w[2] / f[8]);
                    if (zp <= near_ || far_ <= zp) continue;
                    const int pix = yi * is + xi;
                    if (zp < zb[pix] || (zp == zb[pix] && fi[pix] >= 0 && fn < fi[pix])) {
                        zb[pix] = zp;
                        fi[pix] = fn;
                    }
                }
            }
        }
    }
}

/* face_index [B,is,is] -> alpha_full [B,is,is] (unflipped, as the autograd Function
 * returns it) and images [B,R,R] after vertical flip (+ 2x2 average pool when aa). */
void nmr_alpha_flip_pool(const int32_t *face_index, int B, int is, int aa, float *alpha_full, float *images) {
    const int R = aa ? is / 2 : is;
    #pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const int32_t *fi = face_index + (size_t)b * is * is;
        float *af = alpha_full + (size_t)b * is * is;
        float *im = images + (size_t)b * R * R;
        for (int i = 0; i < is * is; ++i) af[i] = fi[i] >= 0 ? 1.f : 0.f;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < R; ++c) {
                if (!aa) {
                    im[r * R + c] = af[(is - 1 - r) * is + c];
                } else {
                    /* flipped rows 2r, 2r+1 are raster rows is-1-2r, is-2-2r; torch avg_pool2d sums then scales */
                    float s = af[(is - 1 - 2 * r) * is + 2 * c] + af[(is - 1 - 2 * r) * is + 2 * c + 1] +
                              af[(is - 2 - 2 * r) * is + 2 * c] + af[(is - 2 - 2 * r) * is + 2 * c + 1];
                    im[r * R + c] = s * 0.25f;
                }
            }
    }
}

/* Approximate gradient of the loss w.r.t. the xy of every face corner.
 * faces [B,nf,9], face_index/alpha/grad_alpha [B,is,is] (unflipped raster frame), grad_faces [B,nf,9]. */
void nmr_pixel_map_bwd(const float *faces, const int32_t *face_index, const float *alpha,
                       const float *grad_alpha, int B, int nf, int is, float eps, float *grad_faces) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long)B * nf; ++i) {
        const int bn = (int)(i / nf), fn = (int)(i % nf);
        const float *f = faces + (size_t)i * 9;
        float g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float *out = grad_faces + (size_t)i * 9;
        if (is_backface(f)) { memset(out, 0, 9 * sizeof(float)); continue; }
        const int32_t *fi = face_index + (size_t)bn * is * is;
        const float *am = alpha + (size_t)bn * is * is;
        const float *gm = grad_alpha + (size_t)bn * is * is;
        for (int e = 0; e < 3; ++e) {
            int pi[3];
            float pp[3][2];
            for (int k = 0; k < 3; ++k) pi[k] = (e + k) % 3;
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 2; ++d) pp[k][d] = 0.5f * (f[3 * pi[k] + d] * is + is - 1);
            for (int axis = 0; axis < 2; ++axis) {
                float p[3][2];
                for (int k = 0; k < 3; ++k)
                    for (int d = 0; d < 2; ++d) p[k][d] = pp[k][(d + axis) % 2];
                int dir;
                if (axis == 0) dir = (p[0][0] < p[1][0]) ? -1 : 1;
                else dir = (p[0][0] < p[1][0]) ? 1 : -1;
                const int d0_from = f2i(fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f));
                const int d0_to = f2i(fminf(fmaxf(p[0][0], p[1][0]), (float)(is - 1)));
                const int slot0 = pi[0] * 3 + (1 - axis), slot1 = pi[1] * 3 + (1 - axis);
                const int stride = (axis == 0) ? is : 1; /* step of +1 in d1 */
                for (int d0 = d0_from; d0 <= d0_to; ++d0) {
                    const float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                    const int d1_in = f2i(dir > 0 ? floorf(d1_cross) : ceilf(d1_cross));
                    const int d1_out = d1_in + dir;
                    if (d1_in < 0 || is <= d1_in) continue;
                    if (d1_out < 0 || is <= d1_out) continue;
                    const int base = (axis == 0) ? d0 : d0 * is; /* index of d1 = 0 on this scan-line */
                    const float alpha_in = am[base + d1_in * stride];
                    const float alpha_out = am[base + d1_out * stride];
                    const float ka = p[1][0] - p[0][0];
                    /* out-sweep: from the out pixel to the image border */
                    if (fi[base + d1_in * stride] == fn) {
                        const int lim = dir > 0 ? is - 1 : 0;
                        int a = d1_out < lim ? d1_out : lim, c = d1_out > lim ? d1_out : lim;
                        if (a < 0) a = 0; if (c > is - 1) c = is - 1;
                        for (int d1 = a; d1 <= c; ++d1) {
                            const float diff = (am[base + d1 * stride] - alpha_in) * gm[base + d1 * stride];
                            if (diff <= 0) continue;
                            if (p[1][0] != d0) {
                                float dist = ka / (p[1][0] - d0) * (d1 - d1_cross) * 2.f / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                g[slot0] -= diff / dist;
                            }
                            if (p[0][0] != d0) {
                                float dist = ka / (d0 - p[0][0]) * (d1 - d1_cross) * 2.f / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                g[slot1] -= diff / dist;
                            }
                        }
                    }
                    /* in-sweep: from the in pixel to the opposite edge of the triangle */
                    {
                        float c2;
                        if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0)
                            c2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                        else
                            c2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
                        const int lim = f2i(dir > 0 ? ceilf(c2) : floorf(c2));
                        int a = d1_in < lim ? d1_in : lim, c = d1_in > lim ? d1_in : lim;
                        if (a < 0) a = 0; if (c > is - 1) c = is - 1;
                        for (int d1 = a; d1 <= c; ++d1) {
                            const float diff = (am[base + d1 * stride] - alpha_out) * gm[base + d1 * stride];
                            if (diff <= 0) continue;
                            if (p[1][0] != d0) {
                                float dist = ka / (p[1][0] - d0) * (d1 - d1_cross) * 2.f / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                g[slot0] -= diff / dist;
                            }
                            if (p[0][0] != d0) {
                                float dist = ka / (d0 - p[0][0]) * (d1 - d1_cross) * 2.f / is;
                                dist = (0 < dist) ? dist + eps : dist - eps;
                                g[slot1] -= diff / dist;
                            }
                        }
                    }
                }
            }
        }
        memcpy(out, g, 9 * sizeof(float));
    }
}
