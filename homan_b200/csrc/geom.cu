// Object rigid placement, vertex-space losses (smoothness, 2-D vertex re-projection, interaction,
// PCA prior), per-problem loss reduction, fused Adam and argmin-over-inits for sm_100a.
//
// Reference code replaced (file:line relative to the reference root):
//   HOMan.get_verts_object          homan/homan.py:298-307
//   compute_smooth_loss             homan/lossutils.py:18-36
//   compute_pca_loss                homan/lossutils.py:39-40
//   compute_verts2d_loss_hand       homan/losses.py:141-164
//   compute_interaction_loss (+ assign_interaction_pairs, project_bbox, compute_iou, compute_dist_z)
//                                   homan/losses.py:20-49,98-139,199-242; homan/utils/bbox.py:115-135;
//                                   homan/utils/geometry.py:69-86
//   loss weighting + Adam groups    homan/jointopt.py:128-151,178-192
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int NVH = HM_MANO_NV;

// ------------------------------------------------------------------------------------------ rigid (object)
__global__ void __launch_bounds__(NT)
rigid_fwd_kernel(const float *__restrict__ mesh, int mesh_batch, const float *__restrict__ rot6d,
                 const float *__restrict__ trans, const float *__restrict__ scale, int V, float *__restrict__ verts) {
    __shared__ Rot6d rs;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) rot6d_forward(rot6d + 6 * b, rs);
    __syncthreads();
    const float s = scale ? fabsf(scale[0]) : 1.f;
    const float t0 = trans[3 * b], t1 = trans[3 * b + 1], t2 = trans[3 * b + 2];
    const float *mb = mesh + (mesh_batch > 1 ? (long)b * V * 3 : 0);
    for (int v = threadIdx.x; v < V; v += NT) {
        const float x = s * mb[3 * v], y = s * mb[3 * v + 1], z = s * mb[3 * v + 2];
        float *o = verts + ((long)b * V + v) * 3;
        o[0] = x * rs.R[0] + y * rs.R[3] + z * rs.R[6] + t0;
        o[1] = x * rs.R[1] + y * rs.R[4] + z * rs.R[7] + t1;
        o[2] = x * rs.R[2] + y * rs.R[5] + z * rs.R[8] + t2;
    }
}

__global__ void __launch_bounds__(NT)
rigid_bwd_kernel(const float *__restrict__ mesh, int mesh_batch, const float *__restrict__ rot6d,
                 const float *__restrict__ scale, int V, const float *__restrict__ g_verts,
                 float *__restrict__ g_rot6d, float *__restrict__ g_trans) {
    __shared__ float red[12 * 32];
    const int b = blockIdx.x;
    const float s = scale ? fabsf(scale[0]) : 1.f;
    const float *mb = mesh + (mesh_batch > 1 ? (long)b * V * 3 : 0);
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    for (int v = threadIdx.x; v < V; v += NT) {
        const float *g = g_verts + ((long)b * V + v) * 3;
        const float g0 = g[0], g1 = g[1], g2 = g[2];
        acc[0] += g0; acc[1] += g1; acc[2] += g2;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float x = s * mb[3 * v + k];
            acc[3 + 3 * k] += x * g0; acc[4 + 3 * k] += x * g1; acc[5 + 3 * k] += x * g2;
        }
    }
    block_sum<12>(acc, red);
    if (threadIdx.x == 0) {
        if (g_trans) for (int c = 0; c < 3; ++c) g_trans[3 * b + c] += acc[c];
        if (g_rot6d) {
            Rot6d rs;
            rot6d_forward(rot6d + 6 * b, rs);
            float g6[6];
            rot6d_backward(rs, acc + 3, g6);
            for (int i = 0; i < 6; ++i) g_rot6d[6 * b + i] += g6[i];
        }
    }
}

// ------------------------------------------------------------------------------------------ vertex losses
// project_bbox of the reference: vertices * (1,-1,1) through nr.projection with orig_size = 1
__device__ __forceinline__ void bbox_project(const float *K, float x, float y, float z, float &u, float &v) {
    const float iz = z + 1e-9f;
    const float xp = x / iz, yp = -y / iz;
    u = K[0] * xp + K[1] * yp + K[2];
    v = K[3] * xp + K[4] * yp + K[5];
    v = 1.f - v;
    u = 2.f * (u - 0.5f);
    v = 2.f * (v - 0.5f);
}

__global__ void __launch_bounds__(NT)
vertex_losses_kernel(const float *__restrict__ vh, const float *__restrict__ vo, const float *__restrict__ camintr,
                     const float *__restrict__ ref2d, const float *__restrict__ pca, int pca_dim, int T, int Vo,
                     float image_size, float w_sh, float w_so, float w_v2d, float w_inter, float w_pca, int flags,
                     float *__restrict__ partials, float *__restrict__ g_vh, float *__restrict__ g_vo,
                     float *__restrict__ g_cdet, float *__restrict__ g_pca) {
    __shared__ float red[12 * 32];
    __shared__ float K[9];
    const int b = blockIdx.x, t = b % T, tid = threadIdx.x;
    if (tid < 9) K[tid] = camintr[9 * b + tid];
    __syncthreads();
    const float *h = vh + (long)b * NVH * 3;
    const float *o = vo + (long)b * Vo * 3;
    float sums[10];  // 0 smooth_h, 1 smooth_o, 2 v2d, 3 v2d px, 4-6 hand centroid, 7-9 obj centroid
#pragma unroll
    for (int i = 0; i < 10; ++i) sums[i] = 0.f;
    float mx[12];    // hand: -umin umax -vmin vmax -zmin zmax ; obj: same
#pragma unroll
    for (int i = 0; i < 12; ++i) mx[i] = -3.4e38f;
    const bool do_smooth = flags & HM_VL_SMOOTH, do_v2d = flags & HM_VL_V2D, do_inter = flags & HM_VL_INTER;
    const bool has_prev = t > 0, has_next = t < T - 1;
    const float n_sh = (float)(T - 1) * NVH * 3.f, n_so = (float)(T - 1) * Vo * 3.f;
    const float n_v2d = (float)T * NVH;

    for (int i = tid; i < NVH; i += NT) {
        const float x = h[3 * i], y = h[3 * i + 1], z = h[3 * i + 2];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (do_smooth) {
            float dn[3] = {0.f, 0.f, 0.f}, dp[3] = {0.f, 0.f, 0.f};
            if (has_next) {
                const float *q = h + NVH * 3 + 3 * i;
                dn[0] = q[0] - x; dn[1] = q[1] - y; dn[2] = q[2] - z;
                sums[0] += dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2];
            }
            if (has_prev) {
                const float *q = h - NVH * 3 + 3 * i;
                dp[0] = x - q[0]; dp[1] = y - q[1]; dp[2] = z - q[2];
            }
            const float c = T > 1 ? w_sh * 2.f / n_sh : 0.f;  // T == 1: no frame pair, no gradient (not inf * 0)
            gx += c * (dp[0] - dn[0]); gy += c * (dp[1] - dn[1]); gz += c * (dp[2] - dn[2]);
        }
        if (do_v2d) {
            const float hx = K[0] * x + K[1] * y + K[2] * z, hy = K[3] * x + K[4] * y + K[5] * z,
                        hz = K[6] * x + K[7] * y + K[8] * z;
            const float px = hx / hz, py = hy / hz;
            const float rx = ref2d[((long)b * NVH + i) * 2], ry = ref2d[((long)b * NVH + i) * 2 + 1];
            const float dx = px - rx / image_size, dy = py - ry / image_size;
            sums[2] += dx * dx + dy * dy;
            const float ex = px * image_size - rx, ey = py * image_size - ry;
            sums[3] += sqrtf(ex * ex + ey * ey);
            const float c = w_v2d * 2.f / n_v2d;
            const float gpx = c * dx, gpy = c * dy;
            const float ghx = gpx / hz, ghy = gpy / hz, ghz = -(gpx * px + gpy * py) / hz;
            gx += K[0] * ghx + K[3] * ghy + K[6] * ghz;
            gy += K[1] * ghx + K[4] * ghy + K[7] * ghz;
            gz += K[2] * ghx + K[5] * ghy + K[8] * ghz;
        }
        if (do_inter) {
            float u, v;
            bbox_project(K, x, y, z, u, v);
            mx[0] = fmaxf(mx[0], -u); mx[1] = fmaxf(mx[1], u); mx[2] = fmaxf(mx[2], -v); mx[3] = fmaxf(mx[3], v);
            mx[4] = fmaxf(mx[4], -z); mx[5] = fmaxf(mx[5], z);
            sums[4] += x; sums[5] += y; sums[6] += z;
        }
        if (g_vh && (do_smooth || do_v2d)) {
            float *g = g_vh + ((long)b * NVH + i) * 3;
            g[0] += gx; g[1] += gy; g[2] += gz;
        }
    }
    for (int i = tid; i < Vo; i += NT) {
        const float x = o[3 * i], y = o[3 * i + 1], z = o[3 * i + 2];
        if (do_smooth) {
            float dn[3] = {0.f, 0.f, 0.f}, dp[3] = {0.f, 0.f, 0.f};
            if (has_next) {
                const float *q = o + (long)Vo * 3 + 3 * i;
                dn[0] = q[0] - x; dn[1] = q[1] - y; dn[2] = q[2] - z;
                sums[1] += dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2];
            }
            if (has_prev) {
                const float *q = o - (long)Vo * 3 + 3 * i;
                dp[0] = x - q[0]; dp[1] = y - q[1]; dp[2] = z - q[2];
            }
            if (g_vo) {
                const float c = T > 1 ? w_so * 2.f / n_so : 0.f;
                float *g = g_vo + ((long)b * Vo + i) * 3;
                g[0] += c * (dp[0] - dn[0]); g[1] += c * (dp[1] - dn[1]); g[2] += c * (dp[2] - dn[2]);
            }
        }
        if (do_inter) {
            float u, v;
            bbox_project(K, x, y, z, u, v);
            mx[6] = fmaxf(mx[6], -u); mx[7] = fmaxf(mx[7], u); mx[8] = fmaxf(mx[8], -v); mx[9] = fmaxf(mx[9], v);
            mx[10] = fmaxf(mx[10], -z); mx[11] = fmaxf(mx[11], z);
            sums[7] += x; sums[8] += y; sums[9] += z;
        }
    }
    block_sum<10>(sums, red);
    if (do_inter) block_max<12>(mx, red);
    float *part = partials + (long)b * HM_NPART;
    if (tid == 0) {
        if (do_smooth) {
            part[HM_PART_SMOOTH_HAND] = T > 1 ? sums[0] / n_sh : 0.f;
            part[HM_PART_SMOOTH_OBJ] = T > 1 ? sums[1] / n_so : 0.f;
        }
        if (do_v2d) {
            part[HM_PART_V2D] = sums[2] / n_v2d;
            part[HM_PART_V2D_PX] = sums[3] / (float)NVH;
        }
        if (do_inter) {
            // boxes (x0, y0, x1, y1) expanded by 0.2 around their centre (losses.py:44-48, expansion 0.2)
            float bh[4], bo[4];
            {
                const float x0 = -mx[0], x1 = mx[1], y0 = -mx[2], y1 = mx[3];
                const float cx = (x0 + x1) / 2.f, cy = (y0 + y1) / 2.f;
                const float ex = (x1 - x0) / 2.f * 1.2f, ey = (y1 - y0) / 2.f * 1.2f;
                bh[0] = cx - ex; bh[1] = cy - ey; bh[2] = cx + ex; bh[3] = cy + ey;
            }
            {
                const float x0 = -mx[6], x1 = mx[7], y0 = -mx[8], y1 = mx[9];
                const float cx = (x0 + x1) / 2.f, cy = (y0 + y1) / 2.f;
                const float ex = (x1 - x0) / 2.f * 1.2f, ey = (y1 - y0) / 2.f * 1.2f;
                bo[0] = cx - ex; bo[1] = cy - ey; bo[2] = cx + ex; bo[3] = cy + ey;
            }
            const float a1 = (bo[2] - bo[0]) * (bo[3] - bo[1]), a2 = (bh[2] - bh[0]) * (bh[3] - bh[1]);
            const float iw = fmaxf(fminf(bo[2], bh[2]) - fmaxf(bo[0], bh[0]), 0.f);
            const float ih = fmaxf(fminf(bo[3], bh[3]) - fmaxf(bo[1], bh[1]), 0.f);
            const float inter = iw * ih;
            const float iou = inter / (a1 + a2 - inter);
            // compute_dist_z(object, hand)
            const float a = -mx[10], bb = mx[11], c = -mx[4], d = mx[5];
            const float zd = (d >= a && bb >= c) ? 0.f : fminf(fabsf(c - bb), fabsf(a - d));
            const bool flag = (iou > 0.f) && (zd < 3.f);
            float loss = 0.f;
            float gc[3] = {0.f, 0.f, 0.f};
            if (flag) {
                for (int k = 0; k < 3; ++k) {
                    const float df = sums[4 + k] / (float)NVH - sums[7 + k] / (float)Vo;
                    loss += df * df;
                    gc[k] = w_inter * 2.f * df / 3.f;
                }
                loss /= 3.f;
            }
            part[HM_PART_INTER] = loss;
            part[HM_PART_INTER_FLAG] = flag ? 1.f : 0.f;
            if (g_cdet) for (int k = 0; k < 3; ++k) g_cdet[3 * b + k] = gc[k];
        }
    }
    if (flags & HM_VL_PCA) {
        float s = 0.f;
        const float n = (float)T * pca_dim;
        for (int k = tid; k < pca_dim; k += NT) {
            const float v = pca[(long)b * pca_dim + k];
            s += v * v;
            if (g_pca) g_pca[(long)b * pca_dim + k] += w_pca * 2.f * v / n;
        }
        float sv[1] = {s};
        block_sum<1>(sv, red);
        if (tid == 0) part[HM_PART_PCA] = sv[0] / n;
    }
}

// ------------------------------------------------------------------------------------------ reduction, Adam
__global__ void finalize_losses_kernel(const float *__restrict__ partials, const float *__restrict__ w, int P, int T,
                                       float *__restrict__ losses, float *__restrict__ total, int *step_counter) {
    const int p = blockIdx.x * blockDim.y + threadIdx.y, k = threadIdx.x;  // blockDim.x == HM_NPART
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && step_counter) step_counter[0] += 1;
    const bool valid = p < P;
    float a = 0.f;
    if (valid) {
        const bool is_mean = k == HM_PART_V2D_PX || k == HM_PART_IOU_OBJ || k == HM_PART_IOU_HAND;
        if (k == HM_PART_MINDIST) {
            a = -3.4e38f;
            for (int t = 0; t < T; ++t) a = fmaxf(a, partials[((long)p * T + t) * HM_NPART + k]);
        } else {
            for (int t = 0; t < T; ++t) a += partials[((long)p * T + t) * HM_NPART + k];
            if (is_mean) a /= (float)T;
        }
        losses[(long)p * HM_NPART + k] = a;
    }
    // weighted total over the 16 slots of this problem (lanes k = 0..15 of one half-warp)
    const float wk = w[k];
    float wa = (valid && wk != 0.f) ? wk * a : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) wa += __shfl_xor_sync(0xffffffffu, wa, o, 16);
    if (valid && k == 0 && total) total[p] = wa;
}

__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, const float *__restrict__ lr, int n, float beta1, float beta2,
                            float eps, const int *__restrict__ step_counter) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float l = lr[i];
    if (l == 0.f) return;
    const int step = step_counter[0];
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)l / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
}

__global__ void argmin_kernel(const float *__restrict__ total, int C, int I, int32_t *__restrict__ best_index,
                              float *__restrict__ best_loss) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float best = INFINITY;
    int bi = 0;
    for (int i = 0; i < I; ++i) {
        const float v = total[(long)c * I + i];
        if (v < best) { best = v; bi = i; }
    }
    best_index[c] = bi;
    if (best_loss) best_loss[c] = best;
}

}  // namespace

// ------------------------------------------------------------------------------------------ object-pose initialiser
// PoseOptimizer.compute_offscreen_loss (homan/pose_optimization.py:112-134) on projected vertices [u, v, z]:
// sum over vertices of relu(u - 1) + relu(v - 1) + relu(-1 - u) + relu(-1 - v) + relu(-z) + relu(z - far).
// One CTA per image; the gradient (weight * d/d ndc, 0 at ties as torch.max(x, 0) of torch 1.6) is accumulated.
__global__ void __launch_bounds__(NT)
offscreen_kernel(const float *__restrict__ ndc, int V, float far_, float weight, float *__restrict__ partials,
                 float *__restrict__ grad_ndc) {
    __shared__ float scratch[32];
    const int b = blockIdx.x;
    const float *p = ndc + (long)b * V * 3;
    float *g = grad_ndc ? grad_ndc + (long)b * V * 3 : nullptr;
    float acc[1] = {0.f};
    for (int i = threadIdx.x; i < V; i += NT) {
        const float u = p[3 * i], v = p[3 * i + 1], z = p[3 * i + 2];
        acc[0] += fmaxf(u - 1.f, 0.f) + fmaxf(v - 1.f, 0.f) + fmaxf(-1.f - u, 0.f) + fmaxf(-1.f - v, 0.f) +
                  fmaxf(-z, 0.f) + fmaxf(z - far_, 0.f);
        if (g) {
            const float gu = (u > 1.f ? 1.f : 0.f) - (u < -1.f ? 1.f : 0.f);
            const float gv = (v > 1.f ? 1.f : 0.f) - (v < -1.f ? 1.f : 0.f);
            const float gz = (z > far_ ? 1.f : 0.f) - (z < 0.f ? 1.f : 0.f);
            if (gu != 0.f) g[3 * i] += weight * gu;
            if (gv != 0.f) g[3 * i + 1] += weight * gv;
            if (gz != 0.f) g[3 * i + 2] += weight * gz;
        }
    }
    block_sum<1>(acc, scratch);
    if (threadIdx.x == 0) partials[(long)b * HM_NPART + HM_PART_OFFSCREEN] = acc[0];
}

// Best candidate ever seen by find_optimal_pose (homan/pose_optimization.py:349-353): when the smallest loss of
// this iteration beats the record, the record takes that loss and the CURRENT parameters of that candidate
// (the reference reads them after optimizer.step()). best = {loss, rot6d[6], trans[3]}. One CTA.
__global__ void __launch_bounds__(NT)
track_best_kernel(const float *__restrict__ total, int N, const float *__restrict__ rot6d,
                  const float *__restrict__ trans, float *__restrict__ best, int32_t *__restrict__ best_index) {
    __shared__ float sv[NT];
    __shared__ int si[NT];
    float v = INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += NT) {
        const float t = total[i];
        if (t < v) { v = t; idx = i; }
    }
    sv[threadIdx.x] = v; si[threadIdx.x] = idx;
    __syncthreads();
    for (int o = NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float ov = sv[threadIdx.x + o];
            const int oi = si[threadIdx.x + o];
            if (ov < sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (sv[0] < best[0] && si[0] < N) {  // (block-uniform)
        const int w = si[0];
        if (threadIdx.x == 0) { best[0] = sv[0]; if (best_index) best_index[0] = w; }
        if (threadIdx.x < 6) best[1 + threadIdx.x] = rot6d[(long)w * 6 + threadIdx.x];
        else if (threadIdx.x < 9) best[1 + threadIdx.x] = trans[(long)w * 3 + threadIdx.x - 6];
    }
}

// Sums of hm_accumulate's fixed-point mode back to float: dst[i] += fixed[i] * 2^-44.
__global__ void fold_fixed_kernel(const unsigned long long *__restrict__ fixed, int n, float *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += __ll2float_rn((long long)fixed[i]) * (1.f / HM_FIXED_SCALE);
}

extern "C" {

int hm_rigid_fwd(const float *mesh, int mesh_batch, const float *rot6d, const float *trans, const float *scale,
                 int B, int V, float *verts, void *stream) {
    HM_NVTX("hm_rigid_fwd");
    HM_REQUIRE(mesh && rot6d && trans && verts, "hm_rigid_fwd: null pointer");
    HM_REQUIRE(B >= 0 && V > 0 && (mesh_batch == 1 || mesh_batch == B), "hm_rigid_fwd: bad sizes");
    if (B == 0) return HM_OK;
    rigid_fwd_kernel<<<B, NT, 0, hm_stream(stream)>>>(mesh, mesh_batch, rot6d, trans, scale, V, verts);
    HM_CHECK_LAUNCH("hm_rigid_fwd");
    return HM_OK;
}

int hm_rigid_bwd(const float *mesh, int mesh_batch, const float *rot6d, const float *scale, int B, int V,
                 const float *grad_verts, float *grad_rot6d, float *grad_trans, void *stream) {
    HM_NVTX("hm_rigid_bwd");
    HM_REQUIRE(mesh && rot6d && grad_verts, "hm_rigid_bwd: null pointer");
    HM_REQUIRE(B >= 0 && V > 0 && (mesh_batch == 1 || mesh_batch == B), "hm_rigid_bwd: bad sizes");
    if (B == 0) return HM_OK;
    rigid_bwd_kernel<<<B, NT, 0, hm_stream(stream)>>>(mesh, mesh_batch, rot6d, scale, V, grad_verts, grad_rot6d,
                                                      grad_trans);
    HM_CHECK_LAUNCH("hm_rigid_bwd");
    return HM_OK;
}

int hm_vertex_losses(const float *verts_hand, const float *verts_obj, const float *camintr,
                     const float *ref_verts2d, const float *pca, int pca_dim, int B, int T, int Vo,
                     float image_size, float w_smooth_hand, float w_smooth_obj, float w_v2d, float w_inter,
                     float w_pca, int flags, float *partials, float *grad_verts_hand, float *grad_verts_obj,
                     float *grad_centroid_det, float *grad_pca, void *stream) {
    HM_NVTX("hm_vertex_losses");
    HM_REQUIRE(verts_hand && verts_obj && camintr && partials, "hm_vertex_losses: null pointer");
    HM_REQUIRE(B >= 0 && T > 0 && B % T == 0 && Vo > 0, "hm_vertex_losses: bad sizes (B must be P*T)");
    HM_REQUIRE(!(flags & HM_VL_V2D) || ref_verts2d, "hm_vertex_losses: v2d needs ref_verts2d");
    HM_REQUIRE(!(flags & HM_VL_PCA) || (pca && pca_dim > 0), "hm_vertex_losses: pca term needs pca");
    if (B == 0) return HM_OK;
    vertex_losses_kernel<<<B, NT, 0, hm_stream(stream)>>>(verts_hand, verts_obj, camintr, ref_verts2d, pca, pca_dim, T,
                                                          Vo, image_size, w_smooth_hand, w_smooth_obj, w_v2d, w_inter,
                                                          w_pca, flags, partials, grad_verts_hand, grad_verts_obj,
                                                          grad_centroid_det, grad_pca);
    HM_CHECK_LAUNCH("hm_vertex_losses");
    return HM_OK;
}

int hm_finalize_losses(const float *partials, const float *weights_part, int P, int T, float *losses,
                       float *total, int *step_counter, void *stream) {
    HM_NVTX("hm_finalize_losses");
    HM_REQUIRE(partials && weights_part && losses, "hm_finalize_losses: null pointer");
    HM_REQUIRE(P >= 0 && T > 0, "hm_finalize_losses: bad sizes");
    if (P == 0) return HM_OK;
    dim3 block(HM_NPART, 8);
    finalize_losses_kernel<<<(P + 7) / 8, block, 0, hm_stream(stream)>>>(partials, weights_part, P, T, losses, total,
                                                                        step_counter);
    HM_CHECK_LAUNCH("hm_finalize_losses");
    return HM_OK;
}

int hm_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                 const float *lr_per_elem, int n, float beta1, float beta2, float eps,
                 const int *step_counter, void *stream) {
    HM_NVTX("hm_adam_step");
    HM_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr_per_elem && step_counter, "hm_adam_step: null pointer");
    HM_REQUIRE(n >= 0, "hm_adam_step: bad size");
    if (n == 0) return HM_OK;
    adam_kernel<<<(n + 255) / 256, 256, 0, hm_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, lr_per_elem, n,
                                                                beta1, beta2, eps, step_counter);
    HM_CHECK_LAUNCH("hm_adam_step");
    return HM_OK;
}

int hm_argmin_over_inits(const float *total, int C, int I, int32_t *best_index, float *best_loss,
                         void *stream) {
    HM_NVTX("hm_argmin_over_inits");
    HM_REQUIRE(total && best_index, "hm_argmin_over_inits: null pointer");
    HM_REQUIRE(C >= 0 && I > 0, "hm_argmin_over_inits: bad sizes");
    if (C == 0) return HM_OK;
    argmin_kernel<<<(C + 127) / 128, 128, 0, hm_stream(stream)>>>(total, C, I, best_index, best_loss);
    HM_CHECK_LAUNCH("hm_argmin_over_inits");
    return HM_OK;
}

int hm_offscreen_loss_fwd_bwd(const float *ndc, int B, int V, float far_, float weight, float *partials,
                              float *grad_ndc, void *stream) {
    HM_NVTX("hm_offscreen_loss_fwd_bwd");
    HM_REQUIRE(ndc && partials, "hm_offscreen_loss_fwd_bwd: null pointer");
    HM_REQUIRE(B >= 0 && V >= 0, "hm_offscreen_loss_fwd_bwd: bad sizes");
    if (B == 0) return HM_OK;
    offscreen_kernel<<<B, NT, 0, hm_stream(stream)>>>(ndc, V, far_, weight, partials, grad_ndc);
    HM_CHECK_LAUNCH("hm_offscreen_loss_fwd_bwd");
    return HM_OK;
}
int hm_track_best(const float *total, int N, const float *rot6d, const float *trans, float *best,
                  int32_t *best_index, void *stream) {
    HM_NVTX("hm_track_best");
    HM_REQUIRE(total && rot6d && trans && best, "hm_track_best: null pointer");
    HM_REQUIRE(N >= 0, "hm_track_best: bad size");
    if (N == 0) return HM_OK;
    track_best_kernel<<<1, NT, 0, hm_stream(stream)>>>(total, N, rot6d, trans, best, best_index);
    HM_CHECK_LAUNCH("hm_track_best");
    return HM_OK;
}

int hm_fold_fixed(const unsigned long long *fixed, int n, float *dst, void *stream) {
    HM_NVTX("hm_fold_fixed");
    HM_REQUIRE(n >= 0, "hm_fold_fixed: bad size");
    if (n == 0) return HM_OK;
    HM_REQUIRE(fixed && dst, "hm_fold_fixed: null pointer");
    fold_fixed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, hm_stream(stream)>>>(fixed, n, dst);
    HM_CHECK_LAUNCH("hm_fold_fixed");
    return HM_OK;
}
}  // extern "C"
