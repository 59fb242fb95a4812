// COMPARATOR, not product code: a scalar CUDA re-creation of the Kato-style `neural_renderer` silhouette
// kernels the reference runs on its GPU path (un-vendored package, README.md:54-58 of the reference; semantics
// as recorded in SURVEY.md Appendix A.3), written the way the upstream extension is organised:
//   forward_face_index_map : one thread per face for the inverse matrices, then one thread per PIXEL looping
//                            over ALL faces, writing face_index, weight and depth maps (24 B / pixel)
//   backward_pixel_map     : one thread per (image, face), three edges x two axes, scan-line sweeps to the
//                            image border / across the triangle, reading full-resolution alpha maps
// It exists to put a number on "the reference's own neural_renderer GPU path" on the same B200
// (scripts/bench_raster_vs_nmr_style.py). Built into baseline/nmr_style/libnmr_style.so; never imported by
// homan_b200.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ bool is_backface(const float *f) {
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

__global__ void face_inv_kernel(const float *__restrict__ faces, int n, int is, float *__restrict__ inv_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *f = faces + (long)i * 9;
    float p[3][2];
    for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 2; ++d) p[k][d] = 0.5f * (f[3 * k + d] * is + is - 1);
    float inv[9] = {p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                    p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                    p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
    const float den = p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]);
    for (int k = 0; k < 9; ++k) inv_out[(long)i * 9 + k] = inv[k] / den;
}

__global__ void face_index_map_kernel(const float *__restrict__ faces, const float *__restrict__ face_inv, int B,
                                      int nf, int is, float near_, float far_, int32_t *__restrict__ face_index,
                                      float *__restrict__ weight_map, float *__restrict__ depth_map) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * is * is) return;
    const int b = (int)(i / ((long)is * is)), pn = (int)(i % ((long)is * is));
    const int yi = pn / is, xi = pn % is;
    const float yp = (float)(2 * yi + 1 - is) / is, xp = (float)(2 * xi + 1 - is) / is;
    const float *fb = faces + (long)b * nf * 9;
    const float *ib = face_inv + (long)b * nf * 9;
    float depth_min = far_, wmin[3] = {0.f, 0.f, 0.f};
    int fmin = -1;
    for (int fn = 0; fn < nf; ++fn) {
        const float *f = fb + (long)fn * 9;
        if (is_backface(f)) continue;
        if ((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) continue;
        if ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) continue;
        if ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])) continue;
        const float *inv = ib + (long)fn * 9;
        float w[3], ws = 0.f;
        for (int k = 0; k < 3; ++k) {
            w[k] = inv[3 * k] * xi + inv[3 * k + 1] * yi + inv[3 * k + 2];
            w[k] = fminf(fmaxf(w[k], 0.f), 1.f);
            ws += w[k];
        }
        for (int k = 0; k < 3; ++k) w[k] /= ws;
        const float zp = 1.f / (w[0] / f[2] + w[1] / f[5] + w[2] / f[8]);
        if (zp <= near_ || far_ <= zp) continue;
        if (zp < depth_min) {
            depth_min = zp; fmin = fn;
            wmin[0] = w[0]; wmin[1] = w[1]; wmin[2] = w[2];
        }
    }
    face_index[i] = fmin;
    depth_map[i] = depth_min;
    weight_map[3 * i] = wmin[0]; weight_map[3 * i + 1] = wmin[1]; weight_map[3 * i + 2] = wmin[2];
}

// The "fast" forward of the multiperson fork the reference installs: one thread per
// (image, FACE) walking the face's pixel bounding box, the nearest face of a pixel kept under a per-pixel lock. The
// lock is a 64-bit atomicMin on (depth bits, face) here: the same winner as the upstream first-minimal-face rule,
// deterministic, and no slower than a spin-lock. A second pass unpacks the keys into the three maps upstream writes.
__global__ void face_index_map_face_parallel_kernel(const float *__restrict__ faces, const float *__restrict__ face_inv,
                                                    int B, int nf, int is, float near_, float far_,
                                                    unsigned long long *__restrict__ keys) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * nf) return;
    const int b = (int)(i / nf), fn = (int)(i % nf);
    const float *f = faces + i * 9;
    if (is_backface(f)) return;
    const float *inv = face_inv + i * 9;
    float px[3], py[3];
    for (int k = 0; k < 3; ++k) { px[k] = 0.5f * (f[3 * k] * is + is - 1); py[k] = 0.5f * (f[3 * k + 1] * is + is - 1); }
    const float xmin = fminf(fminf(px[0], px[1]), px[2]), xmax = fmaxf(fmaxf(px[0], px[1]), px[2]);
    const float ymin = fminf(fminf(py[0], py[1]), py[2]), ymax = fmaxf(fmaxf(py[0], py[1]), py[2]);
    if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) return;
    const int x0 = max(__float2int_rd(xmin) - 1, 0), x1 = min(__float2int_ru(xmax) + 1, is - 1);
    const int y0 = max(__float2int_rd(ymin) - 1, 0), y1 = min(__float2int_ru(ymax) + 1, is - 1);
    unsigned long long *kb = keys + (long)b * is * is;
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
            atomicMin(kb + (long)yi * is + xi, ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)fn);
        }
    }
}
__global__ void unpack_keys_kernel(const float *__restrict__ faces, const float *__restrict__ face_inv, int B, int nf,
                                   int is, float far_, const unsigned long long *__restrict__ keys,
                                   int32_t *__restrict__ face_index, float *__restrict__ weight_map,
                                   float *__restrict__ depth_map) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * is * is) return;
    const unsigned long long k = keys[i];
    const int fn = (int)(unsigned)k;
    float w[3] = {0.f, 0.f, 0.f}, depth = far_;
    if (fn >= 0) {
        const int b = (int)(i / ((long)is * is)), pn = (int)(i % ((long)is * is));
        const int yi = pn / is, xi = pn % is;
        const float *inv = face_inv + ((long)b * nf + fn) * 9;
        float ws = 0.f;
        for (int q = 0; q < 3; ++q) {
            w[q] = inv[3 * q] * xi + inv[3 * q + 1] * yi + inv[3 * q + 2];
            w[q] = fminf(fmaxf(w[q], 0.f), 1.f);
            ws += w[q];
        }
        for (int q = 0; q < 3; ++q) w[q] /= ws;
        depth = __uint_as_float((unsigned)(k >> 32));
    }
    face_index[i] = fn;
    depth_map[i] = depth;
    weight_map[3 * i] = w[0]; weight_map[3 * i + 1] = w[1]; weight_map[3 * i + 2] = w[2];
}

__global__ void backward_pixel_map_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index,
                                          const float *__restrict__ alpha, const float *__restrict__ grad_alpha, int B,
                                          int nf, int is, float eps, float *__restrict__ grad_faces) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * nf) return;
    const int bn = (int)(i / nf), fn = (int)(i % nf);
    const float *f = faces + i * 9;
    float g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float *out = grad_faces + i * 9;
    if (is_backface(f)) {
        for (int k = 0; k < 9; ++k) out[k] = 0.f;
        return;
    }
    const int32_t *fi = face_index + (long)bn * is * is;
    const float *am = alpha + (long)bn * is * is;
    const float *gm = grad_alpha + (long)bn * is * is;
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
            const int d0_from = __float2int_rz(fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f));
            const int d0_to = __float2int_rz(fminf(fmaxf(p[0][0], p[1][0]), (float)(is - 1)));
            const int slot0 = pi[0] * 3 + (1 - axis), slot1 = pi[1] * 3 + (1 - axis);
            const int stride = (axis == 0) ? is : 1;
            for (int d0 = d0_from; d0 <= d0_to; ++d0) {
                const float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                const int d1_in = __float2int_rz(dir > 0 ? floorf(d1_cross) : ceilf(d1_cross));
                const int d1_out = d1_in + dir;
                if (d1_in < 0 || is <= d1_in) continue;
                if (d1_out < 0 || is <= d1_out) continue;
                const int base = (axis == 0) ? d0 : d0 * is;
                const float alpha_in = am[base + d1_in * stride];
                const float alpha_out = am[base + d1_out * stride];
                const float ka = p[1][0] - p[0][0];
                if (fi[base + d1_in * stride] == fn) {
                    const int lim = dir > 0 ? is - 1 : 0;
                    int a = min(d1_out, lim), c = max(d1_out, lim);
                    a = max(a, 0); c = min(c, is - 1);
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
                {
                    float c2;
                    if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0)
                        c2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                    else
                        c2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
                    const int lim = __float2int_rz(dir > 0 ? ceilf(c2) : floorf(c2));
                    int a = min(d1_in, lim), c = max(d1_in, lim);
                    a = max(a, 0); c = min(c, is - 1);
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
    for (int k = 0; k < 9; ++k) out[k] = g[k];
}

}  // namespace

extern "C" {

int nmrs_forward_face_index_map(const float *faces, float *face_inv, int B, int nf, int is, float near_, float far_,
                                int32_t *face_index, float *weight_map, float *depth_map, void *stream) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long n = (long)B * nf;
    face_inv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(faces, (int)n, is, face_inv);
    const long px = (long)B * is * is;
    face_index_map_kernel<<<(unsigned)((px + 255) / 256), 256, 0, s>>>(faces, face_inv, B, nf, is, near_, far_,
                                                                      face_index, weight_map, depth_map);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}

/* keys: [B, is, is] 64-bit scratch */
int nmrs_forward_face_index_map_fast(const float *faces, float *face_inv, int B, int nf, int is, float near_, float far_,
                                     unsigned long long *keys, int32_t *face_index, float *weight_map, float *depth_map,
                                     void *stream) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long n = (long)B * nf, px = (long)B * is * is;
    face_inv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(faces, (int)n, is, face_inv);
    cudaMemsetAsync(keys, 0xff, (size_t)px * 8, s);   // empty = far depth bits all ones, face -1
    face_index_map_face_parallel_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(faces, face_inv, B, nf, is, near_,
                                                                                     far_, keys);
    unpack_keys_kernel<<<(unsigned)((px + 255) / 256), 256, 0, s>>>(faces, face_inv, B, nf, is, far_, keys, face_index,
                                                                    weight_map, depth_map);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}

int nmrs_backward_pixel_map(const float *faces, const int32_t *face_index, const float *alpha,
                            const float *grad_alpha, int B, int nf, int is, float eps, float *grad_faces,
                            void *stream) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long n = (long)B * nf;
    backward_pixel_map_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(faces, face_index, alpha, grad_alpha, B, nf,
                                                                         is, eps, grad_faces);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
