// Fused MANO forward / backward for sm_100a: PCA pose -> axis-angle -> Rodrigues -> blend shapes ->
// kinematic chain -> linear blend skinning -> + mano_trans -> rigid placement (rot6d, translation, scale).
//
// Replaces, on the reference's hot path,
//   ManoModel.forward_pca                homan/manomodel.py:84-151 (PCA comps, mean, left-hand flips)
//   mano.model.load(...)(...) (smplx lbs) un-vendored `mano` package, called at manomodel.py:119-123,136-140
//   HOMan.get_verts_hand                 homan/homan.py:341-382 (+ mano_trans, compute_transformation_persp)
//   rot6d_to_matrix                      homan/utils/geometry.py:9-27
//   compute_transformation_persp         homan/utils/camera.py:108-139
// and their autograd. One CTA per hand-frame; all intermediates stay in shared memory; the backward
// recomputes the forward instead of saving activations (the model is 1.45 MB and L2-resident).
#include "common.cuh"

namespace {

constexpr int NV = HM_MANO_NV;      // 778
constexpr int NJ = HM_MANO_NJ;      // 16
constexpr int NV3 = NV * 3;         // 2334
constexpr int NPF = 135;            // pose-corrective features
constexpr int NT = 256;

__constant__ int c_parents[NJ] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};

struct Model {
    const float *v_template, *shapedirs, *posedirs, *J_template, *J_shapedirs, *weights, *comps, *mean;
};
__device__ __host__ inline Model model_view(const float *blob) {
    Model m;
    m.v_template = blob + HM_MANO_OFF_VTEMPLATE;
    m.shapedirs = blob + HM_MANO_OFF_SHAPEDIRS;
    m.posedirs = blob + HM_MANO_OFF_POSEDIRS;
    m.J_template = blob + HM_MANO_OFF_JTEMPLATE;
    m.J_shapedirs = blob + HM_MANO_OFF_JSHAPEDIRS;
    m.weights = blob + HM_MANO_OFF_WEIGHTS;
    m.mean = blob + HM_MANO_OFF_MEAN;
    m.comps = blob + HM_MANO_OFF_COMPS;
    return m;
}

struct Shared {
    float theta[48];
    float beta[10];
    float R[NJ][9];    // per-joint local rotations (row-major)
    float J[NJ][3];    // rest joints
    float W[NJ][9];    // world rotations of the chain
    float q[NJ][3];    // posed joints
    float tA[NJ][3];   // q_j - W_j J_j
    float pf[NPF];
    float Rh[9];       // rigid rotation (columns b1 b2 b3), row-major
    float a1[3], a2[3], b1[3], b2[3], u[3];
    float n1, nu;      // |a1|, |u| (clamped)
    float th[3], mt[3];
    float scale;
    float vposed[NV3];
};

__device__ __forceinline__ void rodrigues(const float *r, float *R) {
    const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
    const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    const float dx = r[0] / angle, dy = r[1] / angle, dz = r[2] / angle;
    float s, c;
    sincosf(angle, &s, &c);
    const float oc = 1.f - c;
    const float dd = dx * dx + dy * dy + dz * dz;
    // I + s K + (1 - c) K K, K = skew(d), K K = d d^T - |d|^2 I
    R[0] = 1.f + oc * (dx * dx - dd); R[1] = -s * dz + oc * dx * dy;      R[2] = s * dy + oc * dx * dz;
    R[3] = s * dz + oc * dx * dy;     R[4] = 1.f + oc * (dy * dy - dd);   R[5] = -s * dx + oc * dy * dz;
    R[6] = -s * dy + oc * dx * dz;    R[7] = s * dx + oc * dy * dz;       R[8] = 1.f + oc * (dz * dz - dd);
}

// d loss / d r from G = d loss / d R (row-major)
__device__ __forceinline__ void rodrigues_bwd(const float *r, const float *G, float *gr) {
    const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
    const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    const float ia = 1.f / angle;
    const float d[3] = {r[0] * ia, r[1] * ia, r[2] * ia};
    const float n[3] = {ex * ia, ey * ia, ez * ia};
    float s, c;
    sincosf(angle, &s, &c);
    const float oc = 1.f - c;
    const float K[9] = {0.f, -d[2], d[1], d[2], 0.f, -d[0], -d[1], d[0], 0.f};
    float K2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    float gK_s = 0.f, gK2 = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { gK_s += G[i] * K[i]; gK2 += G[i] * K2[i]; }
    const float g_angle = c * gK_s + s * gK2;
    // M = s G + (1 - c) (G K^T + K^T G)
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float gkt = 0.f, ktg = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                gkt += G[i * 3 + k] * K[j * 3 + k];  // (G K^T)_ij = sum_k G_ik K_jk
                ktg += K[k * 3 + i] * G[k * 3 + j];  // (K^T G)_ij = sum_k K_ki G_kj
            }
            M[i * 3 + j] = s * G[i * 3 + j] + oc * (gkt + ktg);
        }
    const float gd[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
    const float gdr = gd[0] * r[0] + gd[1] * r[1] + gd[2] * r[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) gr[a] = gd[a] * ia - n[a] * gdr * ia * ia + g_angle * n[a];
}

__device__ __forceinline__ void mat3_mul(const float *A, const float *B, float *C) {  // C = A B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// Everything up to the posed template (steps shared by forward and backward).
__device__ void mano_common(Shared &S, const Model &m, int ncomps, int left, const float *pca, const float *rot,
                            const float *betas, const float *mano_trans, const float *rot6d, const float *trans,
                            const float *scale, const float *vposed_in = nullptr, float *vposed_out = nullptr) {
    const int t = threadIdx.x;
    if (t < 45) {
        float h = 0.f;
        for (int k = 0; k < ncomps; ++k) h += pca[k] * m.comps[k * 45 + t];
        if (left && (t % 3) != 0) h = -h;
        S.theta[3 + t] = h + m.mean[t];
    } else if (t < 48) {
        S.theta[t - 45] = rot[t - 45];
    } else if (t < 58) {
        S.beta[t - 48] = betas ? betas[t - 48] : 0.f;
    } else if (t < 61) {
        S.mt[t - 58] = mano_trans ? mano_trans[t - 58] : 0.f;
    } else if (t == 64) {
        S.scale = scale ? scale[0] : 1.f;
        if (rot6d) {
            // rot6d_to_matrix: r viewed as [3,2]; a1 = r[:,0], a2 = r[:,1]
            float a1[3] = {rot6d[0], rot6d[2], rot6d[4]}, a2[3] = {rot6d[1], rot6d[3], rot6d[5]};
            const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
            float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
            const float dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
            float u[3] = {a2[0] - dp * b1[0], a2[1] - dp * b1[1], a2[2] - dp * b1[2]};
            const float nu = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
            float b2[3] = {u[0] / nu, u[1] / nu, u[2] / nu};
            float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
            for (int i = 0; i < 3; ++i) {
                S.Rh[i * 3] = b1[i]; S.Rh[i * 3 + 1] = b2[i]; S.Rh[i * 3 + 2] = b3[i];
                S.a1[i] = a1[i]; S.a2[i] = a2[i]; S.b1[i] = b1[i]; S.b2[i] = b2[i]; S.u[i] = u[i];
                S.th[i] = trans ? trans[i] : 0.f;
            }
            S.n1 = n1; S.nu = nu;
        }
    }
    __syncthreads();
    if (t < NJ) rodrigues(&S.theta[3 * t], S.R[t]);
    else if (t >= 32 && t < 32 + NJ * 3) {
        const int i = t - 32;
        float j = m.J_template[i];
#pragma unroll
        for (int l = 0; l < 10; ++l) j += m.J_shapedirs[i * 10 + l] * S.beta[l];
        S.J[i / 3][i % 3] = j;
    }
    __syncthreads();
    if (t == 0) {
        for (int j = 0; j < NJ; ++j) {
            const int p = c_parents[j];
            if (p < 0) {
                for (int i = 0; i < 9; ++i) S.W[0][i] = S.R[0][i];
                for (int i = 0; i < 3; ++i) S.q[0][i] = S.J[0][i];
            } else {
                mat3_mul(S.W[p], S.R[j], S.W[j]);
                float d[3] = {S.J[j][0] - S.J[p][0], S.J[j][1] - S.J[p][1], S.J[j][2] - S.J[p][2]};
                for (int i = 0; i < 3; ++i)
                    S.q[j][i] = S.W[p][i * 3] * d[0] + S.W[p][i * 3 + 1] * d[1] + S.W[p][i * 3 + 2] * d[2] + S.q[p][i];
            }
            for (int i = 0; i < 3; ++i)
                S.tA[j][i] = S.q[j][i] - (S.W[j][i * 3] * S.J[j][0] + S.W[j][i * 3 + 1] * S.J[j][1] + S.W[j][i * 3 + 2] * S.J[j][2]);
        }
    } else if (t >= 32 && t < 32 + NPF) {
        const int k = t - 32;
        const int j = 1 + k / 9, e = k % 9;
        S.pf[k] = S.R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    __syncthreads();
    if (vposed_in) {   // the posed template kept by the forward of this iteration: no second pass over posedirs
        for (int i = t; i < NV3; i += NT) S.vposed[i] = __ldg(vposed_in + i);
    } else {
        for (int i = t; i < NV3; i += NT) {
            float v = m.v_template[i];
#pragma unroll
            for (int l = 0; l < 10; ++l) v += m.shapedirs[i * 10 + l] * S.beta[l];
            float acc = 0.f;
#pragma unroll 5
            for (int k = 0; k < NPF; ++k) acc += S.pf[k] * __ldg(m.posedirs + (long)k * NV3 + i);
            S.vposed[i] = v + acc;
            if (vposed_out) vposed_out[i] = v + acc;
        }
    }
    __syncthreads();
}

// blended transform of vertex v: Tm[0..8] rotation part (row-major), Tm[9..11] translation part
__device__ __forceinline__ void blend(const Shared &S, const Model &m, int v, float *Tm) {
    const float4 *w4 = reinterpret_cast<const float4 *>(m.weights + (long)v * NJ);
    float w[NJ];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 x = __ldg(w4 + k);
        w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) Tm[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Tm[i] += w[j] * S.W[j][i];
#pragma unroll
        for (int i = 0; i < 3; ++i) Tm[9 + i] += w[j] * S.tA[j][i];
    }
}

__global__ void __launch_bounds__(NT)
mano_fwd_kernel(const float *__restrict__ blob, int ncomps, int left, const float *__restrict__ pca, int pca_stride,
                const float *__restrict__ rot, const float *__restrict__ betas, const float *__restrict__ mano_trans,
                const float *__restrict__ rot6d, const float *__restrict__ trans, const float *__restrict__ scale,
                float *__restrict__ verts, float *__restrict__ joints, float *__restrict__ vposed) {
    __shared__ Shared S;
    const int b = blockIdx.x;
    const Model m = model_view(blob);
    mano_common(S, m, ncomps, left, pca + (long)b * pca_stride, rot + 3 * b, betas ? betas + 10 * b : nullptr,
                mano_trans ? mano_trans + 3 * b : nullptr, rot6d ? rot6d + 6 * b : nullptr,
                trans ? trans + 3 * b : nullptr, scale, nullptr, vposed ? vposed + (long)b * NV3 : nullptr);
    for (int v = threadIdx.x; v < NV; v += NT) {
        float Tm[12];
        blend(S, m, v, Tm);
        const float p0 = S.vposed[3 * v], p1 = S.vposed[3 * v + 1], p2 = S.vposed[3 * v + 2];
        float x[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = Tm[i * 3] * p0 + Tm[i * 3 + 1] * p1 + Tm[i * 3 + 2] * p2 + Tm[9 + i] + S.mt[i];
        float *o = verts + ((long)b * NV + v) * 3;
        if (rot6d) {
            const float s = S.scale;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                o[c] = s * x[0] * S.Rh[c] + s * x[1] * S.Rh[3 + c] + s * x[2] * S.Rh[6 + c] + S.th[c];
        } else {
            o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
        }
    }
    if (joints && threadIdx.x < NJ * 3) {
        const int j = threadIdx.x / 3, i = threadIdx.x % 3;
        joints[((long)b * NJ + j) * 3 + i] = S.q[j][i] + S.mt[i];
    }
}

struct SharedBwd {
    float gvm[NV3];      // d loss / d (mano-frame vertex)
    float gvp[NV3];      // d loss / d v_posed
    float GW[NJ][9];
    float Gq[NJ][3];
    float gR[NJ][9];
    float gJ[NJ][3];
    float gpf[NPF];
    float gtheta[48];
    float red[16 * 32];
};

__global__ void __launch_bounds__(NT)
mano_bwd_kernel(const float *__restrict__ blob, int ncomps, int left, const float *__restrict__ pca, int pca_stride,
                const float *__restrict__ rot, const float *__restrict__ betas, const float *__restrict__ mano_trans,
                const float *__restrict__ rot6d, const float *__restrict__ trans, const float *__restrict__ scale,
                const float *__restrict__ vposed, const float *__restrict__ g_verts,
                const float *__restrict__ g_centroid_det, float *__restrict__ g_pca, float *__restrict__ g_rot, float *__restrict__ g_betas,
                float *__restrict__ g_mano_trans, float *__restrict__ g_rot6d, float *__restrict__ g_trans) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &S = *reinterpret_cast<Shared *>(smem_raw);
    SharedBwd &Q = *reinterpret_cast<SharedBwd *>(smem_raw + ((sizeof(Shared) + 15) / 16) * 16);
    const int b = blockIdx.x, t = threadIdx.x;
    const Model m = model_view(blob);
    mano_common(S, m, ncomps, left, pca + (long)b * pca_stride, rot + 3 * b, betas ? betas + 10 * b : nullptr,
                mano_trans ? mano_trans + 3 * b : nullptr, rot6d ? rot6d + 6 * b : nullptr,
                trans ? trans + 3 * b : nullptr, scale, vposed ? vposed + (long)b * NV3 : nullptr, nullptr);
    const bool rigid = rot6d != nullptr;
    const float s = rigid ? S.scale : 1.f;
    // ---- pass 1 over vertices: rigid backward, d/d v_posed
    float acc[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = 0.f;  // 0-2 g_t, 3-11 g_Rh, 12-14 sum v_m
    for (int v = t; v < NV; v += NT) {
        float Tm[12];
        blend(S, m, v, Tm);
        const float p0 = S.vposed[3 * v], p1 = S.vposed[3 * v + 1], p2 = S.vposed[3 * v + 2];
        float x[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = Tm[i * 3] * p0 + Tm[i * 3 + 1] * p1 + Tm[i * 3 + 2] * p2 + Tm[9 + i] + S.mt[i];
        const float *gp = g_verts + ((long)b * NV + v) * 3;
        const float g[3] = {gp[0], gp[1], gp[2]};
        float gm[3];
        if (rigid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] += g[c];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[3 + k * 3 + c] += s * x[k] * g[c];
                gm[k] = s * (S.Rh[k * 3] * g[0] + S.Rh[k * 3 + 1] * g[1] + S.Rh[k * 3 + 2] * g[2]);
                acc[12 + k] += x[k];
            }
        } else {
            gm[0] = g[0]; gm[1] = g[1]; gm[2] = g[2];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Q.gvm[3 * v + k] = gm[k];
            Q.gvp[3 * v + k] = Tm[k] * gm[0] + Tm[3 + k] * gm[1] + Tm[6 + k] * gm[2];  // Tm_rot^T gm
        }
    }
    block_sum<15>(acc, Q.red);
    // ---- rigid parameter gradients (thread 0), including the mesh-detached centroid path
    if (t == 0 && rigid) {
        float gt[3] = {acc[0], acc[1], acc[2]};
        float G[9];
        for (int i = 0; i < 9; ++i) G[i] = acc[3 + i];
        if (g_centroid_det) {
            const float *gc = g_centroid_det + 3 * b;
            for (int c = 0; c < 3; ++c) gt[c] += gc[c];
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) G[k * 3 + c] += s * (acc[12 + k] / (float)NV) * gc[c];
        }
        if (g_trans) for (int c = 0; c < 3; ++c) g_trans[3 * b + c] += gt[c];
        if (g_rot6d) {
            float gb1[3] = {G[0], G[3], G[6]}, gb2[3] = {G[1], G[4], G[7]}, gb3[3] = {G[2], G[5], G[8]};
            const float *b1 = S.b1, *b2 = S.b2, *a2 = S.a2;
            // b3 = b1 x b2
            gb1[0] += b2[1] * gb3[2] - b2[2] * gb3[1]; gb1[1] += b2[2] * gb3[0] - b2[0] * gb3[2]; gb1[2] += b2[0] * gb3[1] - b2[1] * gb3[0];
            gb2[0] += gb3[1] * b1[2] - gb3[2] * b1[1]; gb2[1] += gb3[2] * b1[0] - gb3[0] * b1[2]; gb2[2] += gb3[0] * b1[1] - gb3[1] * b1[0];
            // b2 = u / |u|
            const float d2 = gb2[0] * b2[0] + gb2[1] * b2[1] + gb2[2] * b2[2];
            float gu[3] = {(gb2[0] - d2 * b2[0]) / S.nu, (gb2[1] - d2 * b2[1]) / S.nu, (gb2[2] - d2 * b2[2]) / S.nu};
            // u = a2 - (b1.a2) b1
            const float dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
            const float gub1 = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
            float ga2[3], ga1[3];
            for (int i = 0; i < 3; ++i) {
                ga2[i] = gu[i] - gub1 * b1[i];
                gb1[i] += -dp * gu[i] - gub1 * a2[i];
            }
            const float d1 = gb1[0] * b1[0] + gb1[1] * b1[1] + gb1[2] * b1[2];
            for (int i = 0; i < 3; ++i) ga1[i] = (gb1[i] - d1 * b1[i]) / S.n1;
            for (int i = 0; i < 3; ++i) {
                g_rot6d[6 * b + 2 * i] += ga1[i];
                g_rot6d[6 * b + 2 * i + 1] += ga2[i];
            }
        }
    }
    // sum over vertices of gvm -> mano_trans gradient (threads 32..34), after gvm is complete
    __syncthreads();
    if (g_mano_trans && t >= 32 && t < 35) {
        float a = 0.f;
        for (int v = 0; v < NV; ++v) a += Q.gvm[3 * v + (t - 32)];
        g_mano_trans[3 * b + (t - 32)] += a;
    }
    // ---- per-joint accumulators: GW[j] = sum_v w_vj gvm_v (x) (vp_v - J_j), Gq[j] = sum_v w_vj gvm_v
    if (t < NJ * 12) {
        const int j = t / 12, k = t % 12;
        float a = 0.f;
        if (k < 9) {
            const int r = k / 3, c = k % 3;
            const float Jc = S.J[j][c];
            for (int v = 0; v < NV; ++v) a += __ldg(m.weights + v * NJ + j) * Q.gvm[3 * v + r] * (S.vposed[3 * v + c] - Jc);
            Q.GW[j][k] = a;
        } else {
            const int r = k - 9;
            for (int v = 0; v < NV; ++v) a += __ldg(m.weights + v * NJ + j) * Q.gvm[3 * v + r];
            Q.Gq[j][r] = a;
        }
    }
    // ---- pose-corrective path: gpf[k] = sum_i posedirs[k][i] gvp[i]  (one warp per feature)
    {
        const int lane = t & 31, warp = t >> 5;
        for (int k = warp; k < NPF; k += NT / 32) {
            float a = 0.f;
            for (int i = lane; i < NV3; i += 32) a += __ldg(m.posedirs + (long)k * NV3 + i) * Q.gvp[i];
            a = warp_sum(a);
            if (lane == 0) Q.gpf[k] = a;
        }
    }
    __syncthreads();
    // ---- kinematic chain backward (serial, 16 joints)
    if (t == 0) {
        for (int j = 0; j < NJ; ++j)
            for (int i = 0; i < 3; ++i)
                Q.gJ[j][i] = -(S.W[j][i] * Q.Gq[j][0] + S.W[j][3 + i] * Q.Gq[j][1] + S.W[j][6 + i] * Q.Gq[j][2]);
        for (int j = NJ - 1; j >= 1; --j) {
            const int p = c_parents[j];
            const float d[3] = {S.J[j][0] - S.J[p][0], S.J[j][1] - S.J[p][1], S.J[j][2] - S.J[p][2]};
            for (int r = 0; r < 3; ++r) {
                Q.Gq[p][r] += Q.Gq[j][r];
                for (int c = 0; c < 3; ++c) Q.GW[p][r * 3 + c] += Q.Gq[j][r] * d[c];
            }
            for (int i = 0; i < 3; ++i) {
                const float tt = S.W[p][i] * Q.Gq[j][0] + S.W[p][3 + i] * Q.Gq[j][1] + S.W[p][6 + i] * Q.Gq[j][2];
                Q.gJ[j][i] += tt;
                Q.gJ[p][i] -= tt;
            }
            // W_j = W_p R_j : gW_p += gW_j R_j^T ; gR_j = W_p^T gW_j
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    float a = 0.f, e = 0.f;
                    for (int k = 0; k < 3; ++k) {
                        a += Q.GW[j][r * 3 + k] * S.R[j][c * 3 + k];
                        e += S.W[p][k * 3 + r] * Q.GW[j][k * 3 + c];
                    }
                    Q.GW[p][r * 3 + c] += a;
                    Q.gR[j][r * 3 + c] = e + Q.gpf[(j - 1) * 9 + r * 3 + c];
                }
        }
        for (int i = 0; i < 9; ++i) Q.gR[0][i] = Q.GW[0][i];
        for (int i = 0; i < 3; ++i) Q.gJ[0][i] += Q.Gq[0][i];
    }
    __syncthreads();
    // ---- shape gradient: g_beta[l] = sum_i shapedirs[i][l] gvp[i] + sum_{j,c} J_shapedirs[j][c][l] gJ[j][c]
    if (g_betas) {
        float gb[10];
#pragma unroll
        for (int l = 0; l < 10; ++l) gb[l] = 0.f;
        for (int i = t; i < NV3; i += NT) {
            const float g = Q.gvp[i];
#pragma unroll
            for (int l = 0; l < 10; ++l) gb[l] += m.shapedirs[i * 10 + l] * g;
        }
        if (t < NJ * 3) {
            const float g = Q.gJ[t / 3][t % 3];
#pragma unroll
            for (int l = 0; l < 10; ++l) gb[l] += m.J_shapedirs[t * 10 + l] * g;
        }
        block_sum<10>(gb, Q.red);
        if (t < 10) g_betas[10 * b + t] += gb[t];
    }
    // ---- Rodrigues backward, PCA projection
    if (t < NJ) rodrigues_bwd(&S.theta[3 * t], Q.gR[t], &Q.gtheta[3 * t]);
    __syncthreads();
    if (t < 3) {
        if (g_rot) g_rot[3 * b + t] += Q.gtheta[t];
    } else if (t >= 32 && t < 32 + ncomps && g_pca) {
        const int k = t - 32;
        float a = 0.f;
        for (int i = 0; i < 45; ++i) {
            const float g = (left && (i % 3) != 0) ? -Q.gtheta[3 + i] : Q.gtheta[3 + i];
            a += m.comps[k * 45 + i] * g;
        }
        g_pca[(long)b * pca_stride + k] += a;
    }
}

}  // namespace

extern "C" {

int hm_mano_fwd(const float *model, int ncomps, int left, const float *pca, int pca_stride, const float *rot,
                const float *betas, const float *mano_trans, const float *rot6d, const float *trans,
                const float *scale, int B, float *verts, float *joints, float *vposed, void *stream) {
    HM_NVTX("hm_mano_fwd");
    HM_REQUIRE(model && pca && rot && verts, "hm_mano_fwd: null pointer");
    HM_REQUIRE(B >= 0 && ncomps >= 0 && ncomps <= 45 && pca_stride >= ncomps, "hm_mano_fwd: bad sizes");
    if (B == 0) return HM_OK;
    mano_fwd_kernel<<<B, NT, 0, hm_stream(stream)>>>(model, ncomps, left, pca, pca_stride, rot, betas, mano_trans,
                                                     rot6d, trans, scale, verts, joints, vposed);
    HM_CHECK_LAUNCH("hm_mano_fwd");
    return HM_OK;
}

int hm_mano_bwd(const float *model, int ncomps, int left, const float *pca, int pca_stride, const float *rot,
                const float *betas, const float *mano_trans, const float *rot6d, const float *trans,
                const float *scale, int B, const float *vposed, const float *grad_verts,
                const float *grad_centroid_det, float *grad_pca, float *grad_rot, float *grad_betas,
                float *grad_mano_trans, float *grad_rot6d, float *grad_trans, void *stream) {
    HM_NVTX("hm_mano_bwd");
    HM_REQUIRE(model && pca && rot && grad_verts, "hm_mano_bwd: null pointer");
    HM_REQUIRE(B >= 0 && ncomps >= 0 && ncomps <= 45 && pca_stride >= ncomps, "hm_mano_bwd: bad sizes");
    if (B == 0) return HM_OK;
    const size_t smem = ((sizeof(Shared) + 15) / 16) * 16 + sizeof(SharedBwd);
    static HmSmemOptIn opt_in;
    if (int rc = hm_smem_opt_in(mano_bwd_kernel, smem, opt_in, "hm_mano_bwd")) return rc;
    mano_bwd_kernel<<<B, NT, smem, hm_stream(stream)>>>(model, ncomps, left, pca, pca_stride, rot, betas, mano_trans,
                                                        rot6d, trans, scale, vposed, grad_verts, grad_centroid_det, grad_pca,
                                                        grad_rot, grad_betas, grad_mano_trans, grad_rot6d, grad_trans);
    HM_CHECK_LAUNCH("hm_mano_bwd");
    return HM_OK;
}

}  // extern "C"
