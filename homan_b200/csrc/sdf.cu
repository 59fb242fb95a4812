// Sparse signed-distance interpenetration loss for sm_100a.
//
// Replaces, on the reference's hot path, SDFSceneLoss.forward (homan/interactions/scenesdf.py:77-148)
// as called by compute_collision_loss (homan/lossutils.py:43-64): per object an axis-aligned bbox cube
// (centre, half-size = 0.6 * largest extent), phi = clamp(sdf.SDF(faces, normalised verts), 0) on a 32^3
// grid (un-vendored `sdf` CUDA extension: min point-triangle distance, +x ray parity for the sign), then
// F.grid_sample (trilinear, zeros padding, align_corners=False) of phi at the other object's vertices.
//
// The reference evaluates all 32768 voxels x all faces; grid_sample only ever reads the <= 8 voxels
// around each sample, so this kernel (one CTA per image and ordered pair) evaluates
//   1. which voxels are touched (bitmask in shared memory),
//   2. the ray parity of the touched voxel rows (a +x ray is shared by the 32 voxels of a row),
//   3. the min triangle distance of the touched voxels that are inside (phi = 0 elsewhere after the clamp),
//   4. the trilinear samples, their sum and the gradient w.r.t. the sampled vertices.
// Compiled with -fmad=false: the inside/outside predicate uses the same individually rounded fp32
// operations as the CPU oracle (a flipped parity changes phi from 0 to a distance).
#include "common.cuh"

namespace {

constexpr int NT = 256;   // dense grid kernel
constexpr int PNT = 512;  // pair kernel: one CTA per image, long serial phases -> more warps per CTA
constexpr int VLIST = 8192;  // compacted inside voxels per pass
constexpr int G = 32;  // grid size of the reference (scenesdf.py:14)

__device__ __forceinline__ float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// squared distance from p to triangle (a, b, c): closest-point region walk
__device__ float point_tri_dist2(const float *p, const float *a, const float *b, const float *c) {
    float ab[3], ac[3], ap[3], bp[3], cp[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return dot3(ap, ap);
#pragma unroll
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return dot3(bp, bp);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = ap[k] - v * ab[k];
        return dot3(q, q);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return dot3(cp, cp);
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float w = d2 / (d2 - d6);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = ap[k] - w * ac[k];
        return dot3(q, q);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
        for (int k = 0; k < 3; ++k) q[k] = bp[k] - w * (c[k] - b[k]);
        return dot3(q, q);
    }
    const float denom = 1.f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k] = ap[k] - (ab[k] * v + ac[k] * w);
    return dot3(q, q);
}

// Does the ray (py, pz) + s * (+x) pierce triangle (a, b, c)? On a hit returns the x of the crossing.
__device__ __forceinline__ bool ray_x_cross(float py, float pz, const float *a, const float *b, const float *c,
                                            float &xs) {
    const float w0 = (c[1] - b[1]) * (pz - b[2]) - (c[2] - b[2]) * (py - b[1]);
    const float w1 = (a[1] - c[1]) * (pz - c[2]) - (a[2] - c[2]) * (py - c[1]);
    const float w2 = (b[1] - a[1]) * (pz - a[2]) - (b[2] - a[2]) * (py - a[1]);
    const float area = w0 + w1 + w2;
    bool inside;
    if (area > 0.f) inside = (w0 >= 0.f && w1 >= 0.f && w2 >= 0.f);
    else if (area < 0.f) inside = (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
    else return false;
    if (!inside) return false;
    xs = (w0 * a[0] + w1 * b[0] + w2 * c[0]) / area;
    return true;
}

// -1 + (i + 0.5) * 2 / G as the oracle evaluates it; G is a power of two, so "* 2 / G" is one exact scaling
static_assert((G & (G - 1)) == 0, "power-of-two grid");
__device__ __forceinline__ float voxel_centre(int i) { return -1.f + ((float)i + 0.5f) * (2.f / (float)G); }

// bit mask of the voxels i of a row whose centre x is < xs (the ray from those voxels hits at xs)
__device__ __forceinline__ unsigned row_mask_below(float xs) {
    int n = __float2int_rz(ceilf((xs + 1.f) * 16.f - 0.5f));
    n = max(0, min(G, n));
    while (n > 0 && !(voxel_centre(n - 1) < xs)) --n;
    while (n < G && voxel_centre(n) < xs) ++n;
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

struct Corner {
    int x0, y0, z0;
    float fx, fy, fz;  // fractional offsets from corner 0
};
__device__ __forceinline__ Corner unnormalise(float x, float y, float z) {
    Corner c;
    const float ix = ((x + 1.f) * G - 1.f) / 2.f, iy = ((y + 1.f) * G - 1.f) / 2.f, iz = ((z + 1.f) * G - 1.f) / 2.f;
    const float flx = floorf(ix), fly = floorf(iy), flz = floorf(iz);
    // keep far-away samples from overflowing the int conversion; they touch no voxel either way
    c.x0 = __float2int_rz(fminf(fmaxf(flx, -4.f), (float)G + 4.f));
    c.y0 = __float2int_rz(fminf(fmaxf(fly, -4.f), (float)G + 4.f));
    c.z0 = __float2int_rz(fminf(fmaxf(flz, -4.f), (float)G + 4.f));
    c.fx = ix - flx; c.fy = iy - fly; c.fz = iz - flz;
    return c;
}

__global__ void __launch_bounds__(PNT)
sdf_pair_kernel(const float *__restrict__ verts_g, const int32_t *__restrict__ faces, int faces_batch,
                const float *__restrict__ verts_s, int Vg, int Fg, int Vs, float half_factor, float weight,
                float *__restrict__ phi_all, float *__restrict__ partials, float *__restrict__ g_vs,
                float *__restrict__ dist_values) {
    extern __shared__ __align__(16) float lv[];  // Vg * 3 normalised mesh vertices, then Fg bounding spheres
    float4 *sph = reinterpret_cast<float4 *>(lv + ((Vg * 3 + 3) / 4) * 4);
    __shared__ unsigned needed[G * G], inside[G * G];
    __shared__ float red[6 * 32];
    __shared__ float box[4];  // centre xyz, scale
    __shared__ unsigned short vlist[VLIST];
    __shared__ int wtot[PNT / 32];
    __shared__ unsigned short wq[PNT / 32][64];   // per warp: faces waiting for the exact distance test
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *vg = verts_g + (long)b * Vg * 3;
    const float *vs = verts_s + (long)b * Vs * 3;
    float *phi = phi_all + (long)b * G * G * G;
    if (faces_batch > 1) faces += (long)b * Fg * 3;   // one face list per image (clips with different objects)
    // ---- 1. bbox cube of the grid mesh
    float mx[6] = {-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = tid; i < Vg; i += PNT) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = vg[3 * i + k];
            mx[k] = fmaxf(mx[k], -v);
            mx[3 + k] = fmaxf(mx[3 + k], v);
        }
    }
    for (int i = tid; i < G * G; i += PNT) { needed[i] = 0u; inside[i] = 0u; }
    block_max<6>(mx, red);
    if (tid == 0) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float lo = -mx[k], hi = mx[3 + k];
            box[k] = (lo + hi) / 2.f;
            s = fmaxf(s, (hi - lo) * half_factor);
        }
        box[3] = s;
    }
    __syncthreads();
    const float cx = box[0], cy = box[1], cz = box[2], sc = box[3];
    for (int i = tid; i < Vg * 3; i += PNT) lv[i] = (vg[i] - box[i % 3]) / sc;
    // ---- 2. voxels touched by the samples
    for (int i = tid; i < Vs; i += PNT) {
        const Corner c = unnormalise((vs[3 * i] - cx) / sc, (vs[3 * i + 1] - cy) / sc, (vs[3 * i + 2] - cz) / sc);
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int z = c.z0 + dz, y = c.y0 + dy;
                if (z < 0 || z >= G || y < 0 || y >= G) continue;
                unsigned m = 0u;
                if (c.x0 >= 0 && c.x0 < G) m |= 1u << c.x0;
                if (c.x0 + 1 >= 0 && c.x0 + 1 < G) m |= 1u << (c.x0 + 1);
                if (m) atomicOr(&needed[z * G + y], m);
            }
    }
    __syncthreads();
    // bounding sphere of every face (centroid, radius with rounding slack): culls both loops below
    for (int f = tid; f < Fg; f += PNT) {
        const float *a = lv + 3 * __ldg(faces + 3 * f), *bq = lv + 3 * __ldg(faces + 3 * f + 1),
                    *c = lv + 3 * __ldg(faces + 3 * f + 2);
        const float mx_ = (a[0] + bq[0] + c[0]) / 3.f, my_ = (a[1] + bq[1] + c[1]) / 3.f, mz_ = (a[2] + bq[2] + c[2]) / 3.f;
        float r2 = 0.f;
        r2 = fmaxf(r2, (a[0] - mx_) * (a[0] - mx_) + (a[1] - my_) * (a[1] - my_) + (a[2] - mz_) * (a[2] - mz_));
        r2 = fmaxf(r2, (bq[0] - mx_) * (bq[0] - mx_) + (bq[1] - my_) * (bq[1] - my_) + (bq[2] - mz_) * (bq[2] - mz_));
        r2 = fmaxf(r2, (c[0] - mx_) * (c[0] - mx_) + (c[1] - my_) * (c[1] - my_) + (c[2] - mz_) * (c[2] - mz_));
        sph[f] = make_float4(mx_, my_, mz_, sqrtf(r2) * 1.0001f + 1e-6f);
    }
    __syncthreads();
    // ---- 3. ray parity of every touched row. A +x ray from the row (y, z) can only pierce faces whose y-z bounding
    //         box contains the row, so the loop runs over faces (thread per face) and the few rows of their boxes, and
    //         the crossings are folded into the row's parity word with atomicXor (order-independent): ~10 (face, row)
    //         tests per face instead of one sphere test per (touched row, face). The box carries 1e-4 of slack in the
    //         cube's units (the piercing test itself is the oracle's, rounding ~1e-7).
    for (int f = tid; f < Fg; f += PNT) {
        const float *a = lv + 3 * __ldg(faces + 3 * f), *bq = lv + 3 * __ldg(faces + 3 * f + 1),
                    *c = lv + 3 * __ldg(faces + 3 * f + 2);
        const float ymin = fminf(fminf(a[1], bq[1]), c[1]) - 1e-4f, ymax = fmaxf(fmaxf(a[1], bq[1]), c[1]) + 1e-4f;
        const float zmin = fminf(fminf(a[2], bq[2]), c[2]) - 1e-4f, zmax = fmaxf(fmaxf(a[2], bq[2]), c[2]) + 1e-4f;
        // voxel_centre(j) = -1 + (j + 0.5) / 16 in [ymin, ymax]  (clamped before the int conversion)
        const int j0 = __float2int_rz(fminf(fmaxf(ceilf((ymin + 1.f) * (0.5f * G) - 0.5f), 0.f), (float)G));
        const int j1 = __float2int_rz(fminf(fmaxf(floorf((ymax + 1.f) * (0.5f * G) - 0.5f), -1.f), (float)(G - 1)));
        const int k0 = __float2int_rz(fminf(fmaxf(ceilf((zmin + 1.f) * (0.5f * G) - 0.5f), 0.f), (float)G));
        const int k1 = __float2int_rz(fminf(fmaxf(floorf((zmax + 1.f) * (0.5f * G) - 0.5f), -1.f), (float)(G - 1)));
        for (int k = k0; k <= k1; ++k)
            for (int j = j0; j <= j1; ++j) {
                if (!needed[k * G + j]) continue;
                float xs;
                if (ray_x_cross(voxel_centre(j), voxel_centre(k), a, bq, c, xs)) atomicXor(&inside[k * G + j], row_mask_below(xs));
            }
    }
    __syncthreads();
    for (int i = tid; i < G * G; i += PNT) inside[i] &= needed[i];
    __syncthreads();
    // ---- 4. distance of the touched inside voxels: compact them (deterministic prefix sum over the rows),
    //         then one warp per voxel, lanes over faces
    static_assert(G * G == 2 * PNT, "two voxel rows per thread");
    const unsigned bits0 = inside[2 * tid], bits1 = inside[2 * tid + 1];
    const int c0 = __popc(bits0), c1 = __popc(bits1);
    int incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int excl = incl - (c0 + c1);
    for (int w = 0; w < warp; ++w) excl += wtot[w];
    int total = 0;
    for (int w = 0; w < PNT / 32; ++w) total += wtot[w];
    for (int base = 0; base < total; base += VLIST) {
        {
            int g = excl;
            unsigned m = bits0;
            while (m) {
                const int ix = __ffs(m) - 1;
                m &= m - 1;
                if (g >= base && g < base + VLIST) vlist[g - base] = (unsigned short)((2 * tid) * G + ix);
                ++g;
            }
            m = bits1;
            while (m) {
                const int ix = __ffs(m) - 1;
                m &= m - 1;
                if (g >= base && g < base + VLIST) vlist[g - base] = (unsigned short)((2 * tid + 1) * G + ix);
                ++g;
            }
        }
        __syncthreads();
        const int nv = min(total - base, VLIST);
        // a warp takes a contiguous stretch of the list (neighbouring voxels): only its first voxel needs the pass
        // over all bounding spheres for an upper bound of its distance - afterwards the previous voxel's distance plus
        // the step between the two centres is one (triangle inequality), and usually a tighter one
        const int per = (nv + PNT / 32 - 1) / (PNT / 32);
        float prev_d = -1.f, pp[3] = {0.f, 0.f, 0.f};
        for (int i = warp * per; i < min(nv, (warp + 1) * per); ++i) {
            const int vox = vlist[i];
            const int row = vox / G, ix = vox % G;
            const float p[3] = {voxel_centre(ix), voxel_centre(row % G), voxel_centre(row / G)};
            float ub;
            if (prev_d >= 0.f) {
                const float sx = p[0] - pp[0], sy = p[1] - pp[1], sz = p[2] - pp[2];
                ub = prev_d + sqrtf(sx * sx + sy * sy + sz * sz);
            } else {
                ub = INFINITY;
                for (int f = lane; f < Fg; f += 32) {
                    const float4 sp = sph[f];
                    const float dx = p[0] - sp.x, dy = p[1] - sp.y, dz = p[2] - sp.z;
                    ub = fminf(ub, sqrtf(dx * dx + dy * dy + dz * dz) + sp.w);
                }
                ub = warp_min(ub);
            }
            ub = ub * 1.0001f + 1e-6f;
            // faces whose sphere reaches inside the bound are queued per warp and evaluated 32 at a time: the exact
            // point-triangle distance is ~150 divergent instructions and only a few lanes of every 32 spheres pass
            float best = INFINITY;
            int qn = 0;
            unsigned short *q = wq[warp];
            for (int f0 = 0; f0 < Fg; f0 += 32) {
                const int f = f0 + lane;
                bool cand = false;
                if (f < Fg) {
                    const float4 sp = sph[f];
                    const float dx = p[0] - sp.x, dy = p[1] - sp.y, dz = p[2] - sp.z, t = ub + sp.w;
                    cand = !(dx * dx + dy * dy + dz * dz > t * t);
                }
                const unsigned m = __ballot_sync(0xffffffffu, cand);
                if (m == 0u) continue;
                if (cand) q[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)f;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    const int fq = q[lane];
                    const int i0 = __ldg(faces + 3 * fq), i1 = __ldg(faces + 3 * fq + 1), i2 = __ldg(faces + 3 * fq + 2);
                    best = fminf(best, point_tri_dist2(p, lv + 3 * i0, lv + 3 * i1, lv + 3 * i2));
                    const int rest = qn - 32;
                    const int mv = lane < rest ? q[32 + lane] : 0;
                    __syncwarp();
                    if (lane < rest) q[lane] = (unsigned short)mv;
                    qn = rest;
                    __syncwarp();
                }
            }
            if (lane < qn) {
                const int fq = q[lane];
                const int i0 = __ldg(faces + 3 * fq), i1 = __ldg(faces + 3 * fq + 1), i2 = __ldg(faces + 3 * fq + 2);
                best = fminf(best, point_tri_dist2(p, lv + 3 * i0, lv + 3 * i1, lv + 3 * i2));
            }
            __syncwarp();
            best = warp_min(best);
            const float d = sqrtf(best);
            if (lane == 0) phi[vox] = d;
            prev_d = d < INFINITY ? d : -1.f;   // (an empty mesh or non-finite vertices: keep using the sphere pass)
            pp[0] = p[0]; pp[1] = p[1]; pp[2] = p[2];
        }
        __syncthreads();
    }
    // ---- 5. trilinear samples (grid_sample, zeros padding) + gradient w.r.t. the sampled vertices
    float acc[1] = {0.f};
    for (int i = tid; i < Vs; i += PNT) {
        const Corner c = unnormalise((vs[3 * i] - cx) / sc, (vs[3 * i + 1] - cy) / sc, (vs[3 * i + 2] - cz) / sc);
        float out = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int x = c.x0 + dx, y = c.y0 + dy, z = c.z0 + dz;
                    if (x < 0 || x >= G || y < 0 || y >= G || z < 0 || z >= G) continue;
                    if (!((inside[z * G + y] >> x) & 1u)) continue;
                    const float val = phi[(z * G + y) * G + x];
                    const float wx = dx ? c.fx : 1.f - c.fx, wy = dy ? c.fy : 1.f - c.fy, wz = dz ? c.fz : 1.f - c.fz;
                    out += val * wx * wy * wz;
                    gx += val * (dx ? 1.f : -1.f) * wy * wz;
                    gy += val * wx * (dy ? 1.f : -1.f) * wz;
                    gz += val * wx * wy * (dz ? 1.f : -1.f);
                }
        acc[0] += out;
        if (dist_values) dist_values[(long)b * Vs + i] = out * sc;   // back in the scene's units (scenesdf.py:143-146)
        if (g_vs && weight != 0.f) {
            const float k = weight * (0.5f * G) / sc;
            float *g = g_vs + ((long)b * Vs + i) * 3;
            g[0] += k * gx; g[1] += k * gy; g[2] += k * gz;
        }
    }
    block_sum<1>(acc, red);
    if (tid == 0) partials[(long)b * HM_NPART + HM_PART_COLLISION] += acc[0];
}

// Dense grid (drop-in for sdf.SDF): one thread per voxel, all faces.
__global__ void __launch_bounds__(NT)
sdf_grid_kernel(const int32_t *__restrict__ faces, const float *__restrict__ verts, int V, int F, int Gs,
                float *__restrict__ phi) {
    const int b = blockIdx.y;
    const int vox = blockIdx.x * NT + threadIdx.x;
    if (vox >= Gs * Gs * Gs) return;
    const float *vb = verts + (long)b * V * 3;
    const int i = vox % Gs, j = (vox / Gs) % Gs, k = vox / (Gs * Gs);
    const float p[3] = {-1.f + ((float)i + 0.5f) * 2.f / (float)Gs, -1.f + ((float)j + 0.5f) * 2.f / (float)Gs,
                        -1.f + ((float)k + 0.5f) * 2.f / (float)Gs};
    float best = INFINITY;
    int hits = 0;
    for (int f = 0; f < F; ++f) {
        const float *a = vb + 3 * __ldg(faces + 3 * f), *bb = vb + 3 * __ldg(faces + 3 * f + 1),
                    *c = vb + 3 * __ldg(faces + 3 * f + 2);
        const float A[3] = {__ldg(a), __ldg(a + 1), __ldg(a + 2)}, Bv[3] = {__ldg(bb), __ldg(bb + 1), __ldg(bb + 2)},
                    C[3] = {__ldg(c), __ldg(c + 1), __ldg(c + 2)};
        best = fminf(best, point_tri_dist2(p, A, Bv, C));
        float xs;
        if (ray_x_cross(p[1], p[2], A, Bv, C, xs) && xs > p[0]) ++hits;
    }
    const float d = sqrtf(best);
    phi[(long)b * Gs * Gs * Gs + vox] = (hits & 1) ? d : -d;
}

}  // namespace

extern "C" {

int hm_sdf_pair(const float *verts_g, const int32_t *faces_g, int faces_batch, const float *verts_s, int B, int Vg,
                int Fg, int Vs, int grid, float scale_factor, float weight, float *phi_scratch, float *partials,
                float *grad_verts_s, float *dist_values, void *stream) {
    HM_NVTX("hm_sdf_pair");
    HM_REQUIRE(verts_g && faces_g && verts_s && phi_scratch && partials, "hm_sdf_pair: null pointer");
    HM_REQUIRE(B >= 0 && Vg > 0 && Fg > 0 && Vs > 0 && (faces_batch == 1 || faces_batch == B), "hm_sdf_pair: bad sizes");
    HM_UNSUPPORTED(grid != G, "hm_sdf_pair: grid size %d (only %d, the reference's grid_size)", grid, G);
    const size_t smem = (size_t)((Vg * 3 + 3) / 4) * 4 * sizeof(float) + (size_t)Fg * 4 * sizeof(float);
    HM_UNSUPPORTED(smem > 160 * 1024, "hm_sdf_pair: grid mesh with %d vertices does not fit shared memory", Vg);
    if (B == 0) return HM_OK;
    static HmSmemOptIn opt_in;
    if (int rc = hm_smem_opt_in(sdf_pair_kernel, smem, opt_in, "hm_sdf_pair")) return rc;
    const float half_factor = (float)((1.0 + (double)scale_factor) * 0.5);
    sdf_pair_kernel<<<B, PNT, smem, hm_stream(stream)>>>(verts_g, faces_g, faces_batch, verts_s, Vg, Fg, Vs, half_factor,
                                                        weight, phi_scratch, partials, grad_verts_s, dist_values);
    HM_CHECK_LAUNCH("hm_sdf_pair");
    return HM_OK;
}

int hm_sdf_grid(const int32_t *faces, const float *verts, int B, int V, int F, int grid, float *phi,
                void *stream) {
    HM_NVTX("hm_sdf_grid");
    HM_REQUIRE(faces && verts && phi, "hm_sdf_grid: null pointer");
    HM_REQUIRE(B >= 0 && B <= 65535 && V > 0 && F > 0 && grid > 0 && grid <= 256, "hm_sdf_grid: bad sizes");
    if (B == 0) return HM_OK;
    dim3 gridDim((grid * grid * grid + NT - 1) / NT, B);
    sdf_grid_kernel<<<gridDim, NT, 0, hm_stream(stream)>>>(faces, verts, V, F, grid, phi);
    HM_CHECK_LAUNCH("hm_sdf_grid");
    return HM_OK;
}

}  // extern "C"
