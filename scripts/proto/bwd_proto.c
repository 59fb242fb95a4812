/*
 * Development tool (not product, not oracle): scalar CPU model of the enumeration used by the sm_100a raster
 * backward (homan_b200/csrc/raster.cu), to check on the CPU that it visits exactly the crossings of
 * backward_pixel_map (oracle/csrc/nmr_raster.c) and to count the work it does.
 *
 *   pass O  out-sweeps found from the pixels: a span end of the face_index map (owner changes along the sweep
 *           direction) looks up the owner's edges; the crossing whose in-pixel is that pixel (or the one before it)
 *           sweeps the missing-coverage pixels beyond.
 *   pass S  out-sweeps of tasks whose slope is too steep for the span-end argument (|slope| > SMAX or not finite):
 *           enumerated from the face as the reference does.
 *   pass I  in-sweeps: only faces whose pixel bounding box holds an uncovered pixel (or whose corners sit on
 *           integer pixel coordinates) can contribute.
 * Sweeps are summed pixel by pixel here (the closed-form run sums of the kernel are tested elsewhere).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)v;
}
static inline int is_backface(const float *f) { return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]); }

typedef struct {
    float p[3][2];   /* p0 p1 p2 along (d0, d1) */
    int pi[3], dir, d0_from, d0_to;
    float slope, ka;
} Task;

static void make_task(const float *f, int is, int e, int axis, Task *t) {
    float pp[3][2];
    for (int k = 0; k < 3; ++k) t->pi[k] = (e + k) % 3;
    for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 2; ++d) pp[k][d] = 0.5f * (f[3 * t->pi[k] + d] * is + is - 1);
    for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 2; ++d) t->p[k][d] = pp[k][(d + axis) % 2];
    if (axis == 0) t->dir = (t->p[0][0] < t->p[1][0]) ? -1 : 1;
    else t->dir = (t->p[0][0] < t->p[1][0]) ? 1 : -1;
    t->d0_from = f2i(fmaxf(ceilf(fminf(t->p[0][0], t->p[1][0])), 0.f));
    t->d0_to = f2i(fminf(fmaxf(t->p[0][0], t->p[1][0]), (float)(is - 1)));
    t->ka = t->p[1][0] - t->p[0][0];
    t->slope = (t->p[1][1] - t->p[0][1]) / t->ka;
}

static void add_pixel(const Task *t, int d0, int d1, float d1_cross, float diff, int is, float eps, float *g0, float *g1) {
    if (t->p[1][0] != d0) {
        float dist = t->ka / (t->p[1][0] - d0) * (d1 - d1_cross) * 2.f / is;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *g0 -= diff / dist;
    }
    if (t->p[0][0] != d0) {
        float dist = t->ka / (d0 - t->p[0][0]) * (d1 - d1_cross) * 2.f / is;
        dist = (0 < dist) ? dist + eps : dist - eps;
        *g1 -= diff / dist;
    }
}


/* ---- model of the kernel's sweep arithmetic: run lists + closed-form harmonic tails (raster.cu: build_runs_kernel,
 * eval_item, sweep_line). g_runs != 0 switches out_sweep / in-sweeps of proto_pixel_map_bwd to it. */
static int g_runs = 0;
void proto_set_runs(int r) { g_runs = r; }
#define NEAR_N 4
#define RCAP 8
static float harmonic_span(float z1, float n) {
    const float z2 = z1 + n;
    const float i1 = 1.f / z1, i2 = 1.f / z2;
    const float a1 = i1 * i1, a2 = i2 * i2;
    float r = log1pf(n * i1);
    r += 0.5f * (i1 - i2);
    r += (1.f / 12.f) * (a1 - a2);
    r -= (1.f / 120.f) * (a1 * a1 - a2 * a2);
    r += (1.f / 252.f) * (a1 * a1 * a1 - a2 * a2 * a2);
    return r;
}
static void eval_item(float x, float c0, float c1, float G, int s, int e, int has0, int has1, float inv_is2, float eps,
                      float *a0, float *a1) {
    const float K0 = c0 * inv_is2, K1 = c1 * inv_is2;
    const float rK0 = 1.f / K0, rK1 = 1.f / K1;
    const float del0 = eps * fabsf(rK0), del1 = eps * fabsf(rK1);
    const int left = (float)e <= x;
    const float sgn = left ? -1.f : 1.f;
    const float z = left ? x - (float)e : (float)s - x;
    const int n = e - s + 1;
    float h0 = 0.f, h1 = 0.f;
    for (int k = 0; k < NEAR_N; ++k) {
        const float dd = sgn * (z + (float)k);
        float dist0 = K0 * dd, dist1 = K1 * dd;
        dist0 = (0.f < dist0) ? dist0 + eps : dist0 - eps;
        dist1 = (0.f < dist1) ? dist1 + eps : dist1 - eps;
        const float t0 = 1.f / dist0, t1 = 1.f / dist1;
        if (k < n) { h0 += t0; h1 += t1; }
    }
    const float nf = (float)(n - NEAR_N > 0 ? n - NEAR_N : 0), zf = z + (float)NEAR_N;
    h0 += sgn * harmonic_span(zf + del0, nf) * rK0;
    h1 += sgn * harmonic_span(zf + del1, nf) * rK1;
    *a0 = has0 ? -G * h0 : 0.f;
    *a1 = has1 ? -G * h1 : 0.f;
}
/* sweep of the pixels d1 in [ra, rc] of one line with weight w(d1) = max(0, (alpha[d1] - alpha_ref) * g[d1]) */
static void sweep_model(const Task *t, int axis, int d0, float x, int ra, int rc, float alpha_ref, const float *am,
                        const float *gm, int is, float eps, int walk, float *g0, float *g1, long *n_items) {
    const int stride = (axis == 0) ? is : 1, base = (axis == 0) ? d0 : d0 * is;
    const int has0 = t->p[1][0] != d0, has1 = t->p[0][0] != d0;
    const float c0 = t->ka / (t->p[1][0] - d0), c1 = t->ka / (d0 - t->p[0][0]);
    /* runs of the WHOLE line (the kernel's lists are per line, not per sweep) */
    int rs[1024], re[1024], n = 0;
    float rg[1024];
    int cur = -1;
    for (int d1 = 0; d1 < is; ++d1) {
        const float diff = (am[base + d1 * stride] - alpha_ref) * gm[base + d1 * stride];
        if (diff <= 0) { cur = -1; continue; }
        if (cur >= 0 && re[cur] == d1 - 1 && rg[cur] == diff) { re[cur] = d1; continue; }
        cur = n++;
        rs[cur] = re[cur] = d1;
        rg[cur] = diff;
    }
    if (n > RCAP || walk) {
        for (int d1 = ra; d1 <= rc; ++d1) {
            const float diff = (am[base + d1 * stride] - alpha_ref) * gm[base + d1 * stride];
            if (diff <= 0) continue;
            const float dd = (float)d1 - x;
            if (has0) { float dist = c0 * dd * 2.f / (float)is; dist = (0.f < dist) ? dist + eps : dist - eps; *g0 -= diff / dist; }
            if (has1) { float dist = c1 * dd * 2.f / (float)is; dist = (0.f < dist) ? dist + eps : dist - eps; *g1 -= diff / dist; }
        }
        return;
    }
    for (int r = 0; r < n; ++r) {
        const int s = ra > rs[r] ? ra : rs[r], e = rc < re[r] ? rc : re[r];
        if (s > e) continue;
        float a0, a1;
        eval_item(x, c0, c1, rg[r], s, e, has0, has1, 2.f / (float)is, eps, &a0, &a1);
        *g0 += a0;
        *g1 += a1;
        ++*n_items;
    }
}

static void out_sweep(const Task *t, int axis, int d0, float x, int d1_in, int d1_out, const float *am, const float *gm,
                      int is, float eps, float *g, long *n_px) {
    const int stride = (axis == 0) ? is : 1, base = (axis == 0) ? d0 : d0 * is;
    const float alpha_in = am[base + d1_in * stride];
    const int lim = t->dir > 0 ? is - 1 : 0;
    int a = d1_out < lim ? d1_out : lim, c = d1_out > lim ? d1_out : lim;
    if (a < 0) a = 0;
    if (c > is - 1) c = is - 1;
    float g0 = 0, g1 = 0;
    if (g_runs) {
        sweep_model(t, axis, d0, x, a, c, alpha_in, am, gm, is, eps, 0, &g0, &g1, n_px + 4);
    } else
    for (int d1 = a; d1 <= c; ++d1) {
        const float diff = (am[base + d1 * stride] - alpha_in) * gm[base + d1 * stride];
        if (diff <= 0) continue;
        add_pixel(t, d0, d1, x, diff, is, eps, &g0, &g1);
        ++*n_px;
    }
    g[t->pi[0] * 3 + (1 - axis)] += g0;
    g[t->pi[1] * 3 + (1 - axis)] += g1;
}

/* stats: 0 span-end candidates, 1 matched out-sweeps, 2 (candidate, edge) evaluations reaching the slope, 3 steep tasks,
 * 4 steep crossings, 5 boundary faces, 6 front faces, 7 in-sweep crossings, 8 in-sweep crossings with a contribution,
 * 9 out-sweep pixels, 10 in-sweep pixels, 11 matched at q0 - dir, 12 irregular faces */
static int g_mode = 7; /* bit 0: pass O, bit 1: pass S, bit 2: pass I */
void proto_set_mode(int m) { g_mode = m; }

void proto_pixel_map_bwd(const float *faces, const int32_t *face_index, const float *alpha, const float *grad_alpha,
                         int B, int nf, int is, float eps, float smax, float *grad_faces, long *stats) {
    memset(grad_faces, 0, sizeof(float) * (size_t)B * nf * 9);
    for (int b = 0; b < B; ++b) {
        const float *fb = faces + (size_t)b * nf * 9;
        const int32_t *fi = face_index + (size_t)b * is * is;
        const float *am = alpha + (size_t)b * is * is, *gm = grad_alpha + (size_t)b * is * is;
        float *gb = grad_faces + (size_t)b * nf * 9;
        /* ---- pass O */
        for (int yi = 0; yi < is && (g_mode & 1); ++yi)
            for (int xi = 0; xi < is; ++xi) {
                const int fn = fi[yi * is + xi];
                if (fn < 0) continue;
                const float *f = fb + (size_t)fn * 9;
                for (int axis = 0; axis < 2; ++axis) {
                    const int d0 = axis == 0 ? xi : yi, q0 = axis == 0 ? yi : xi;
                    const int stride = (axis == 0) ? is : 1, base = (axis == 0) ? d0 : d0 * is;
                    for (int dir = -1; dir <= 1; dir += 2) {
                        const int qn = q0 + dir;
                        if (qn >= 0 && qn < is && fi[base + qn * stride] == fn) continue; /* not the end of a span */
                        ++stats[0];
                        for (int e = 0; e < 3; ++e) {
                            Task t;
                            make_task(f, is, e, axis, &t);
                            if (t.dir != dir || d0 < t.d0_from || d0 > t.d0_to) continue;
                            ++stats[2];
                            if (!(fabsf(t.slope) <= smax)) continue;
                            const float x = t.slope * (d0 - t.p[0][0]) + t.p[0][1];
                            const int d1_in = f2i(dir > 0 ? floorf(x) : ceilf(x)), d1_out = d1_in + dir;
                            if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                            if (d1_in != q0) {
                                if (d1_in != q0 - dir || fi[base + d1_in * stride] != fn) continue;
                                ++stats[11];
                            }
                            ++stats[1];
                            out_sweep(&t, axis, d0, x, d1_in, d1_out, am, gm, is, eps, gb + (size_t)fn * 9, &stats[9]);
                        }
                    }
                }
            }
        /* ---- pass S and pass I: per face */
        for (int fn = 0; fn < nf; ++fn) {
            const float *f = fb + (size_t)fn * 9;
            if (is_backface(f)) continue;
            ++stats[6];
            float *g = gb + (size_t)fn * 9;
            float pmin[2] = {INFINITY, INFINITY}, pmax[2] = {-INFINITY, -INFINITY};
            int irregular = 0;
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 2; ++d) {
                    const float v = 0.5f * (f[3 * k + d] * is + is - 1);
                    if (!(fabsf(v) < 1e30f) || v == floorf(v)) irregular = 1;
                    pmin[d] = fminf(pmin[d], v);
                    pmax[d] = fmaxf(pmax[d], v);
                }
            int boundary = irregular;
            stats[12] += irregular;
            if (!boundary) {
                int x0 = f2i(floorf(pmin[0])) - 1, x1 = f2i(ceilf(pmax[0])) + 1;
                int y0 = f2i(floorf(pmin[1])) - 1, y1 = f2i(ceilf(pmax[1])) + 1;
                if (x0 < 0) x0 = 0;
                if (y0 < 0) y0 = 0;
                if (x1 > is - 1) x1 = is - 1;
                if (y1 > is - 1) y1 = is - 1;
                for (int y = y0; y <= y1 && !boundary; ++y)
                    for (int x = x0; x <= x1; ++x)
                        if (fi[y * is + x] < 0) { boundary = 1; break; }
            }
            stats[5] += boundary;
            for (int e = 0; e < 3; ++e)
                for (int axis = 0; axis < 2; ++axis) {
                    Task t;
                    make_task(f, is, e, axis, &t);
                    const int steep = !(fabsf(t.slope) <= smax);
                    if (!steep && !boundary) continue;
                    stats[3] += steep;
                    const int stride = (axis == 0) ? is : 1;
                    for (int d0 = t.d0_from; d0 <= t.d0_to; ++d0) {
                        const float x = t.slope * (d0 - t.p[0][0]) + t.p[0][1];
                        const int d1_in = f2i(t.dir > 0 ? floorf(x) : ceilf(x)), d1_out = d1_in + t.dir;
                        if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                        const int base = (axis == 0) ? d0 : d0 * is;
                        if (steep && (g_mode & 2)) {
                            ++stats[4];
                            if (fi[base + d1_in * stride] == fn)
                                out_sweep(&t, axis, d0, x, d1_in, d1_out, am, gm, is, eps, g, &stats[9]);
                        }
                        if (!boundary || !(g_mode & 4)) continue;
                        ++stats[7];
                        const float alpha_out = am[base + d1_out * stride];
                        float c2;
                        if ((d0 - t.p[0][0]) * (d0 - t.p[2][0]) < 0)
                            c2 = (t.p[2][1] - t.p[0][1]) / (t.p[2][0] - t.p[0][0]) * (d0 - t.p[0][0]) + t.p[0][1];
                        else
                            c2 = (t.p[1][1] - t.p[2][1]) / (t.p[1][0] - t.p[2][0]) * (d0 - t.p[2][0]) + t.p[2][1];
                        const int lim = f2i(t.dir > 0 ? ceilf(c2) : floorf(c2));
                        int a = d1_in < lim ? d1_in : lim, c = d1_in > lim ? d1_in : lim;
                        if (a < 0) a = 0;
                        if (c > is - 1) c = is - 1;
                        float g0 = 0, g1 = 0;
                        long n = 0;
                        if (g_runs) {
                            const int walk = (float)a < x - (float)(NEAR_N - 1) && (float)c > x;
                            if (a <= c) sweep_model(&t, axis, d0, x, a, c, alpha_out, am, gm, is, eps, walk, &g0, &g1, &stats[14]);
                            n = g0 != 0 || g1 != 0;
                        } else
                        for (int d1 = a; d1 <= c; ++d1) {
                            const float diff = (am[base + d1 * stride] - alpha_out) * gm[base + d1 * stride];
                            if (diff <= 0) continue;
                            add_pixel(&t, d0, d1, x, diff, is, eps, &g0, &g1);
                            ++n;
                        }
                        stats[10] += n;
                        stats[8] += n > 0;
                        g[t.pi[0] * 3 + (1 - axis)] += g0;
                        g[t.pi[1] * 3 + (1 - axis)] += g1;
                    }
                }
        }
    }
}

/* In-sweep contributions of the faces pass I skips (must be exactly zero): returns the number of non-zero ones. */
long proto_skipped_insweeps(const float *faces, const int32_t *face_index, const float *alpha, const float *grad_alpha,
                            int B, int nf, int is) {
    long bad = 0;
    for (int b = 0; b < B; ++b) {
        const float *fb = faces + (size_t)b * nf * 9;
        const int32_t *fi = face_index + (size_t)b * is * is;
        const float *am = alpha + (size_t)b * is * is, *gm = grad_alpha + (size_t)b * is * is;
        for (int fn = 0; fn < nf; ++fn) {
            const float *f = fb + (size_t)fn * 9;
            if (is_backface(f)) continue;
            float pmin[2] = {INFINITY, INFINITY}, pmax[2] = {-INFINITY, -INFINITY};
            int irregular = 0;
            for (int k = 0; k < 3; ++k)
                for (int d = 0; d < 2; ++d) {
                    const float v = 0.5f * (f[3 * k + d] * is + is - 1);
                    if (!(fabsf(v) < 1e30f) || v == floorf(v)) irregular = 1;
                    pmin[d] = fminf(pmin[d], v);
                    pmax[d] = fmaxf(pmax[d], v);
                }
            if (irregular) continue;
            int x0 = f2i(floorf(pmin[0])) - 1, x1 = f2i(ceilf(pmax[0])) + 1;
            int y0 = f2i(floorf(pmin[1])) - 1, y1 = f2i(ceilf(pmax[1])) + 1;
            if (x0 < 0) x0 = 0;
            if (y0 < 0) y0 = 0;
            if (x1 > is - 1) x1 = is - 1;
            if (y1 > is - 1) y1 = is - 1;
            int boundary = 0;
            for (int y = y0; y <= y1 && !boundary; ++y)
                for (int x = x0; x <= x1; ++x)
                    if (fi[y * is + x] < 0) { boundary = 1; break; }
            if (boundary) continue;
            for (int e = 0; e < 3; ++e)
                for (int axis = 0; axis < 2; ++axis) {
                    Task t;
                    make_task(f, is, e, axis, &t);
                    const int stride = (axis == 0) ? is : 1;
                    for (int d0 = t.d0_from; d0 <= t.d0_to; ++d0) {
                        const float x = t.slope * (d0 - t.p[0][0]) + t.p[0][1];
                        const int d1_in = f2i(t.dir > 0 ? floorf(x) : ceilf(x)), d1_out = d1_in + t.dir;
                        if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                        const int base = (axis == 0) ? d0 : d0 * is;
                        const float alpha_out = am[base + d1_out * stride];
                        float c2;
                        if ((d0 - t.p[0][0]) * (d0 - t.p[2][0]) < 0)
                            c2 = (t.p[2][1] - t.p[0][1]) / (t.p[2][0] - t.p[0][0]) * (d0 - t.p[0][0]) + t.p[0][1];
                        else
                            c2 = (t.p[1][1] - t.p[2][1]) / (t.p[1][0] - t.p[2][0]) * (d0 - t.p[2][0]) + t.p[2][1];
                        const int lim = f2i(t.dir > 0 ? ceilf(c2) : floorf(c2));
                        int a = d1_in < lim ? d1_in : lim, c = d1_in > lim ? d1_in : lim;
                        if (a < 0) a = 0;
                        if (c > is - 1) c = is - 1;
                        for (int d1 = a; d1 <= c; ++d1)
                            if ((am[base + d1 * stride] - alpha_out) * gm[base + d1 * stride] > 0) ++bad;
                    }
                }
        }
    }
    return bad;
}
