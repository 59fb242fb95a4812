// Silhouette rasteriser (forward + approximate backward) and pin-hole projection for sm_100a.
//
// Replaces the `neural_renderer` CUDA extension on the reference's hot path:
//   nr.projection                       -> project_fwd_kernel / project_bwd_kernel
//   fill_back + vertices_to_faces       -> face_setup_kernel (front-facing winding only, no doubling)
//   forward_face_index_map + alpha,
//   vertical flip, 2x2 avg-pool (AA)    -> raster_fwd_kernel (64x64 tile, shared-memory z-buffer)
//   backward_pixel_map + scatter-add    -> grad_prep_kernel + raster_bwd_kernel
// Call sites in the reference: homan/losses.py:34-41,73-77,172-176,187; homan/homan.py:168-176.
//
// The file is compiled with -fmad=false: coverage / ownership predicates are evaluated with the same
// individually rounded fp32 operations as the CPU oracle so that face_index maps agree bit for bit.
#include "common.cuh"

namespace {

constexpr int TILE = 64;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

struct __align__(16) FaceRec {
    float c[9];   // x0 y0 z0 x1 y1 z1 x2 y2 z2 in NDC, front-facing winding
    int fn;       // face number in the doubled (fill_back) numbering, -1 = culled
    int v[3];     // vertex indices in that winding
    short bb[4];  // pixel bbox x0 y0 x1 y1 (clamped), empty when x0 > x1
    int pad;
    float inv[9];  // barycentric matrix in pixel coordinates, already divided by its determinant
    float pad2[7];
};
static_assert(sizeof(FaceRec) == HM_FACE_RECORD_BYTES, "record size");

struct __align__(8) FaceBox {
    short x0, y0, x1, y1;
};
static_assert(sizeof(FaceBox) == HM_FACE_BBOX_BYTES, "bbox size");

// ------------------------------------------------------------------------------------------ projection
__global__ void project_fwd_kernel(const float *__restrict__ verts, const float *__restrict__ K, int K_batch,
                                   const float *__restrict__ R, const float *__restrict__ t,
                                   const float *__restrict__ dist, int dist_batch, float orig_size, float eps,
                                   int B, int V, float *__restrict__ ndc) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * V) return;
    const int b = (int)(i / V);
    float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    if (R) {
        const float X = R[0] * x + R[1] * y + R[2] * z, Y = R[3] * x + R[4] * y + R[5] * z,
                    Z = R[6] * x + R[7] * y + R[8] * z;
        x = X; y = Y; z = Z;
    }
    if (t) { x += t[0]; y += t[1]; z += t[2]; }
    const float x_ = x / (z + eps), y_ = y / (z + eps);
    float x__ = x_, y__ = y_;
    if (dist) {
        const float *d = dist + (dist_batch > 1 ? 5 * b : 0);
        const float k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
        const float r = sqrtf(x_ * x_ + y_ * y_);
        const float r2 = r * r, r4 = r2 * r2, r6 = r4 * r2;
        const float rad = 1.f + k1 * r2 + k2 * r4 + k3 * r6;
        x__ = x_ * rad + 2.f * p1 * x_ * y_ + p2 * (r2 + 2.f * x_ * x_);
        y__ = y_ * rad + p1 * (r2 + 2.f * y_ * y_) + 2.f * p2 * x_ * y_;
    }
    const float *Kb = K + (K_batch > 1 ? 9 * b : 0);
    float u = Kb[0] * x__ + Kb[1] * y__ + Kb[2];
    float v = Kb[3] * x__ + Kb[4] * y__ + Kb[5];
    v = orig_size - v;
    u = 2.f * (u - orig_size / 2.f) / orig_size;
    v = 2.f * (v - orig_size / 2.f) / orig_size;
    ndc[3 * i] = u;
    ndc[3 * i + 1] = v;
    ndc[3 * i + 2] = z;
}

__global__ void project_bwd_kernel(const float *__restrict__ verts, const float *__restrict__ K, int K_batch,
                                   const float *__restrict__ R, const float *__restrict__ t, float orig_size,
                                   float eps, int B, int V, const float *__restrict__ grad_ndc,
                                   float *__restrict__ grad_verts, int accumulate) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * V) return;
    const int b = (int)(i / V);
    float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    if (R) {
        const float X = R[0] * x + R[1] * y + R[2] * z, Y = R[3] * x + R[4] * y + R[5] * z,
                    Z = R[6] * x + R[7] * y + R[8] * z;
        x = X; y = Y; z = Z;
    }
    if (t) { x += t[0]; y += t[1]; z += t[2]; }
    const float iz = 1.f / (z + eps);
    const float x_ = x * iz, y_ = y * iz;
    const float *Kb = K + (K_batch > 1 ? 9 * b : 0);
    const float gu = grad_ndc[3 * i] * (2.f / orig_size), gv = -grad_ndc[3 * i + 1] * (2.f / orig_size);
    const float gx_ = gu * Kb[0] + gv * Kb[3], gy_ = gu * Kb[1] + gv * Kb[4];
    float gX = gx_ * iz, gY = gy_ * iz, gZ = -(gx_ * x_ + gy_ * y_) * iz + grad_ndc[3 * i + 2];
    if (R) {
        const float a = R[0] * gX + R[3] * gY + R[6] * gZ, c = R[1] * gX + R[4] * gY + R[7] * gZ,
                    d = R[2] * gX + R[5] * gY + R[8] * gZ;
        gX = a; gY = c; gZ = d;
    }
    if (accumulate) {
        grad_verts[3 * i] += gX; grad_verts[3 * i + 1] += gY; grad_verts[3 * i + 2] += gZ;
    } else {
        grad_verts[3 * i] = gX; grad_verts[3 * i + 1] = gY; grad_verts[3 * i + 2] = gZ;
    }
}

// ------------------------------------------------------------------------------------------ face setup
__device__ __forceinline__ float to_pix(float v, int is) { return 0.5f * (v * is + is - 1); }

__global__ void face_setup_kernel(const float *__restrict__ ndc, const int32_t *__restrict__ faces, int faces_batch,
                                  int B, int V, int F, int is, int fill_back, FaceRec *__restrict__ recs,
                                  FaceBox *__restrict__ boxes) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * F) return;
    const int b = (int)(i / F), f = (int)(i % F);
    const int32_t *fc = faces + ((faces_batch > 1 ? (long)b * F : 0) + f) * 3;
    int vi[3] = {fc[0], fc[1], fc[2]};
    float p[3][3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (vi[k] < 0 || vi[k] >= V) { ok = false; vi[k] = 0; }
        const float *q = ndc + ((long)b * V + vi[k]) * 3;
        p[k][0] = q[0]; p[k][1] = q[1]; p[k][2] = q[2];
    }
    int fn = -1;
    bool rev = false;
    // back-face test of the original winding, then of the reversed copy appended by fill_back
    if (!((p[2][1] - p[0][1]) * (p[1][0] - p[0][0]) < (p[1][1] - p[0][1]) * (p[2][0] - p[0][0]))) {
        fn = f;
    } else if (fill_back &&
               !((p[0][1] - p[2][1]) * (p[1][0] - p[2][0]) < (p[1][1] - p[2][1]) * (p[0][0] - p[2][0]))) {
        fn = F + f;
        rev = true;
    }
    if (!ok) fn = -1;
    FaceRec r;
    const int o0 = rev ? 2 : 0, o2 = rev ? 0 : 2;
    const int ord[3] = {o0, 1, o2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.c[3 * k] = p[ord[k]][0]; r.c[3 * k + 1] = p[ord[k]][1]; r.c[3 * k + 2] = p[ord[k]][2];
        r.v[k] = vi[ord[k]];
    }
    // conservative pixel bbox, 1 px of slack (covers every in/out pixel of backward_pixel_map too)
    const float px0 = to_pix(p[0][0], is), px1 = to_pix(p[1][0], is), px2 = to_pix(p[2][0], is);
    const float py0 = to_pix(p[0][1], is), py1 = to_pix(p[1][1], is), py2 = to_pix(p[2][1], is);
    const float xmin = fminf(fminf(px0, px1), px2), xmax = fmaxf(fmaxf(px0, px1), px2);
    const float ymin = fminf(fminf(py0, py1), py2), ymax = fmaxf(fmaxf(py0, py1), py2);
    if (!(px0 == px0) || !(px1 == px1) || !(px2 == px2) || !(py0 == py0) || !(py1 == py1) || !(py2 == py2)) fn = -1;
    int x0 = __float2int_rz(floorf(xmin)) - 1, x1 = __float2int_rz(ceilf(xmax)) + 1;
    int y0 = __float2int_rz(floorf(ymin)) - 1, y1 = __float2int_rz(ceilf(ymax)) + 1;
    if (fn < 0 || x1 < 0 || y1 < 0 || x0 > is - 1 || y0 > is - 1) {
        x0 = 1; x1 = 0; y0 = 1; y1 = 0;  // empty
        if (fn >= 0) fn = -2 - fn;       // off-screen: no pixel, no crossing (kept negative => skipped)
    } else {
        x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, is - 1); y1 = min(y1, is - 1);
    }
    r.fn = fn;
    r.bb[0] = (short)x0; r.bb[1] = (short)y0; r.bb[2] = (short)x1; r.bb[3] = (short)y1;
    r.pad = 0;
    {   // barycentric matrix of the stored winding (same expressions as the oracle's per-face setup)
        const float p00 = to_pix(r.c[0], is), p01 = to_pix(r.c[1], is), p10 = to_pix(r.c[3], is), p11 = to_pix(r.c[4], is),
                    p20 = to_pix(r.c[6], is), p21 = to_pix(r.c[7], is);
        const float m[9] = {p11 - p21, p20 - p10, p10 * p21 - p20 * p11,
                            p21 - p01, p00 - p20, p20 * p01 - p00 * p21,
                            p01 - p11, p10 - p00, p00 * p11 - p10 * p01};
        const float den = p20 * (p01 - p11) + p00 * (p11 - p21) + p10 * (p21 - p01);
#pragma unroll
        for (int k = 0; k < 9; ++k) r.inv[k] = m[k] / den;
#pragma unroll
        for (int k = 0; k < 7; ++k) r.pad2[k] = 0.f;
    }
    recs[i] = r;
    FaceBox bx;
    bx.x0 = (short)x0; bx.y0 = (short)y0; bx.x1 = (short)x1; bx.y1 = (short)y1;
    boxes[i] = bx;
}

constexpr int LISTCAP = 2048;  // faces of one tile processed per batch

constexpr int SCAN = 4 * NTHREADS;  // faces tested per scan step (four 8-byte boxes per thread)

// Appends to list[*cnt...] the faces of [base, base + SCAN) whose bbox touches the tile.
__device__ __forceinline__ void append_faces(const FaceBox *__restrict__ boxes, int base, int F, int tx0, int ty0,
                                             int *list, int *cnt) {
    const int f0 = base + 4 * threadIdx.x;
    FaceBox bx[4];
    if (f0 + 3 < F && (reinterpret_cast<uintptr_t>(boxes + f0) & 15u) == 0) {  // two 16-byte loads of 4 boxes
        const uint4 u0 = __ldg(reinterpret_cast<const uint4 *>(boxes + f0));
        const uint4 u1 = __ldg(reinterpret_cast<const uint4 *>(boxes + f0) + 1);
        const unsigned w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bx[k].x0 = (short)(w[2 * k] & 0xffff); bx[k].y0 = (short)(w[2 * k] >> 16);
            bx[k].x1 = (short)(w[2 * k + 1] & 0xffff); bx[k].y1 = (short)(w[2 * k + 1] >> 16);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (f0 + k < F) bx[k] = boxes[f0 + k];
            else { bx[k].x0 = 1; bx[k].x1 = 0; bx[k].y0 = 1; bx[k].y1 = 0; }
        }
    }
    const int lane = threadIdx.x & 31;
    unsigned hits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool hit = bx[k].x0 <= bx[k].x1 && bx[k].x0 <= tx0 + TILE - 1 && bx[k].x1 >= tx0 &&
                         bx[k].y0 <= ty0 + TILE - 1 && bx[k].y1 >= ty0;
        hits |= (hit ? 1u : 0u) << k;
    }
    const int mine = __popc(hits);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int wbase = 0;
    if (lane == 31 && total) wbase = atomicAdd(cnt, total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    int pos = wbase + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((hits >> k) & 1u) list[pos++] = f0 + k;
}

// Fills the list with the next batch of faces touching the tile (scanning from *base). Block-uniform.
__device__ __forceinline__ int next_batch(const FaceBox *__restrict__ boxes, int &base, int F, int tx0, int ty0,
                                          int *list, int *cnt, int *next, int cap = LISTCAP) {
    __syncthreads();
    if (threadIdx.x == 0) { *cnt = 0; *next = 0; }
    __syncthreads();
    int n = 0;
    while (base < F && n <= cap - SCAN) {
        append_faces(boxes, base, F, tx0, ty0, list, cnt);
        base += SCAN;
        __syncthreads();
        n = *cnt;
    }
    return n;
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NTHREADS)
raster_fwd_kernel(const FaceRec *__restrict__ recs, const FaceBox *__restrict__ boxes, int F, int is, int aa,
                  float near_, float far_, int32_t *__restrict__ face_index, float *__restrict__ alpha,
                  uint32_t *__restrict__ cov_row, uint32_t *__restrict__ cov_col) {
    __shared__ unsigned long long keys[TILE * TILE];
    __shared__ int list[LISTCAP];
    __shared__ int cnt, next;
    __shared__ uint32_t roww[TILE][2];
    __shared__ unsigned short pixq[NWARPS][64];
    __shared__ short spanx[NWARPS][64], spanpre[NWARPS][64];
    const int b = blockIdx.y;
    const int tiles_x = is / TILE;
    const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
    const int lane = threadIdx.x & 31;
    const bool pow2 = (is & (is - 1)) == 0;
    const float inv_is = 1.f / (float)is;
    const unsigned long long empty = ((unsigned long long)__float_as_uint(far_) << 32) | 0xffffffffull;
    for (int i = threadIdx.x; i < TILE * TILE; i += NTHREADS) keys[i] = empty;
    recs += (long)b * F;
    boxes += (long)b * F;

    int base = 0;
    while (base < F) {
        const int n = next_batch(boxes, base, F, tx0, ty0, list, &cnt, &next);
        // two passes: faces kept in their original winding (the outer layer of an outward-wound closed mesh)
        // first, the reversed copies second, so that early z culls the hidden layer before its depth maths
        for (;;) {  // warps pull (pass, face) pairs off the list
            int li0 = 0;
            if (lane == 0) li0 = atomicAdd(&next, 1);
            li0 = __shfl_sync(0xffffffffu, li0, 0);
            if (li0 >= 2 * n) break;
            const int pass = li0 >= n ? 1 : 0, li = li0 - pass * n;
            const FaceRec *rp = recs + list[li];
            const float4 q0 = __ldg(reinterpret_cast<const float4 *>(rp));
            const float4 q1 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
            const float4 q2 = __ldg(reinterpret_cast<const float4 *>(rp) + 2);
            const int4 q3 = __ldg(reinterpret_cast<const int4 *>(rp) + 3);
            const float f0 = q0.x, f1 = q0.y, f2 = q0.z, f3 = q0.w, f4 = q1.x, f5 = q1.y, f6 = q1.z, f7 = q1.w,
                        f8 = q2.x;
            const int fn = __float_as_int(q2.y);
            if ((fn >= F ? 1 : 0) != pass) continue;
            const int bx0 = (short)(q3.y & 0xffff), by0 = (short)(q3.y >> 16);
            const int bx1 = (short)(q3.z & 0xffff), by1 = (short)(q3.z >> 16);
            const int X0 = max(bx0, tx0), X1 = min(bx1, tx0 + TILE - 1);
            const int Y0 = max(by0, ty0), Y1 = min(by1, ty0 + TILE - 1);
            const int w = X1 - X0 + 1, h = Y1 - Y0 + 1;
            if (w <= 0 || h <= 0) continue;
            const float e0x = f3 - f0, e0y = f4 - f1, e1x = f6 - f3, e1y = f7 - f4, e2x = f0 - f6, e2y = f1 - f7;
            // early z: the interpolated depth is a convex combination of the corner depths, so a face whose
            // nearest corner (minus rounding slack) is behind the current winner of a pixel cannot win it
            const float zmin = fminf(fminf(f2, f5), f8);
            const unsigned zmin_bits = zmin > 0.f ? __float_as_uint(zmin * (1.f - 1e-5f)) : 0u;
            // Candidate pixels: per row of the bbox, the conservative x-span of the triangle (each edge bounds x
            // from one side), flattened over the lanes through a prefix sum of the span lengths - slivers and
            // diagonal faces cost their area, not their bounding box. The exact test below decides coverage.
            short *rowx = spanx[threadIdx.x >> 5], *rowpre = spanpre[threadIdx.x >> 5];
            int n_px;
            {
                int len[2], xlo[2];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r = lane + 32 * half;
                    len[half] = 0; xlo[half] = X0;
                    if (r < h) {
                        const int yi = Y0 + r;
                        const float yp = pow2 ? (float)(2 * yi + 1 - is) * inv_is : (float)(2 * yi + 1 - is) / (float)is;
                        float lo = -3.0e38f, hi = 3.0e38f;
                        bool empty = false;
                        const float c0 = (yp - f1) * e0x, c1 = (yp - f4) * e1x, c2 = (yp - f7) * e2x;
                        if (e0y > 0.f) hi = fminf(hi, f0 + __fdividef(c0, e0y)); else if (e0y < 0.f) lo = fmaxf(lo, f0 + __fdividef(c0, e0y)); else empty |= c0 < -1e-6f;
                        if (e1y > 0.f) hi = fminf(hi, f3 + __fdividef(c1, e1y)); else if (e1y < 0.f) lo = fmaxf(lo, f3 + __fdividef(c1, e1y)); else empty |= c1 < -1e-6f;
                        if (e2y > 0.f) hi = fminf(hi, f6 + __fdividef(c2, e2y)); else if (e2y < 0.f) lo = fmaxf(lo, f6 + __fdividef(c2, e2y)); else empty |= c2 < -1e-6f;
                        // NaN bounds (degenerate edges) fall back to the whole row
                        int a = X0, c = X1;
                        const float plo = to_pix(lo, is), phi = to_pix(hi, is);
                        if (plo == plo) a = max(X0, __float2int_rz(fminf(fmaxf(ceilf(plo) - 1.f, -1.f), 70000.f)));
                        if (phi == phi) c = min(X1, __float2int_rz(fminf(fmaxf(floorf(phi) + 1.f, -1.f), 70000.f)));
                        xlo[half] = a;
                        len[half] = empty ? 0 : max(c - a + 1, 0);
                    }
                }
                int inc0 = len[0], inc1 = len[1];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v0 = __shfl_up_sync(0xffffffffu, inc0, o), v1 = __shfl_up_sync(0xffffffffu, inc1, o);
                    if (lane >= o) { inc0 += v0; inc1 += v1; }
                }
                const int tot0 = __shfl_sync(0xffffffffu, inc0, 31), tot1 = __shfl_sync(0xffffffffu, inc1, 31);
                __syncwarp();
                rowx[lane] = (short)xlo[0]; rowpre[lane] = (short)(inc0 - len[0]);
                rowx[lane + 32] = (short)xlo[1]; rowpre[lane + 32] = (short)(tot0 + inc1 - len[1]);
                n_px = tot0 + tot1;
                __syncwarp();
            }
            if (n_px == 0) continue;  // the bounding box touches the tile, the triangle does not
            // barycentric matrix in pixel coordinates, prepared once per face by the setup kernel
            float inv[9];
            {
                const float4 i0 = __ldg(reinterpret_cast<const float4 *>(rp) + 4);
                const float4 i1 = __ldg(reinterpret_cast<const float4 *>(rp) + 5);
                inv[0] = i0.x; inv[1] = i0.y; inv[2] = i0.z; inv[3] = i0.w;
                inv[4] = i1.x; inv[5] = i1.y; inv[6] = i1.z; inv[7] = i1.w;
                inv[8] = __ldg(reinterpret_cast<const float *>(rp) + 24);
            }
            // Inside pixels are compacted into a per-warp queue so that the depth maths (7 IEEE divisions) runs
            // on full warps.
            unsigned short *pq = pixq[threadIdx.x >> 5];
            int qn = 0;
            for (int i0 = 0; i0 < n_px || qn > 0; i0 += 32) {
                bool in = false;
                unsigned packed = 0;
                const int i = i0 + lane;
                if (i < n_px) {
                    int row = 0;
#pragma unroll
                    for (int sft = 32; sft > 0; sft >>= 1)
                        if (row + sft < h && rowpre[row + sft] <= i) row += sft;
                    const int xi = rowx[row] + (i - rowpre[row]), yi = Y0 + row;
                    // (2i + 1 - is) / is: for a power-of-two raster the division is an exact scaling
                    const float yp = pow2 ? (float)(2 * yi + 1 - is) * inv_is : (float)(2 * yi + 1 - is) / (float)is;
                    const float xp = pow2 ? (float)(2 * xi + 1 - is) * inv_is : (float)(2 * xi + 1 - is) / (float)is;
                    in = !((yp - f1) * e0x < (xp - f0) * e0y) && !((yp - f4) * e1x < (xp - f3) * e1y) &&
                         !((yp - f7) * e2x < (xp - f6) * e2y);
                    packed = (unsigned)(xi - tx0) | ((unsigned)(yi - ty0) << 6);
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) pq[qn + __popc(m & ((1u << lane) - 1u))] = (unsigned short)packed;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32 || (i0 + 32 >= n_px && qn > 0)) {
                    const int take = min(qn, 32);
                    if (lane < take) {
                        const unsigned e = pq[lane];
                        const int xl = e & 63u, yl = e >> 6;
                        const int xi = tx0 + xl, yi = ty0 + yl;
                        unsigned long long *kp = &keys[yl * TILE + xl];
                        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
                        if (!(zmin_bits > (unsigned)(cur >> 32))) {
                            float w0 = inv[0] * xi + inv[1] * yi + inv[2];
                            float w1 = inv[3] * xi + inv[4] * yi + inv[5];
                            float w2 = inv[6] * xi + inv[7] * yi + inv[8];
                            w0 = fminf(fmaxf(w0, 0.f), 1.f);
                            w1 = fminf(fmaxf(w1, 0.f), 1.f);
                            w2 = fminf(fmaxf(w2, 0.f), 1.f);
                            const float ws = w0 + w1 + w2;
                            w0 /= ws; w1 /= ws; w2 /= ws;
                            const float zp = 1.f / (w0 / f2 + w1 / f5 + w2 / f8);
                            if (zp > near_ && zp < far_) {  // also rejects NaN
                                const unsigned long long key = ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)fn;
                                if (key < cur) atomicMin(kp, key);
                            }
                        }
                    }
                    const int rem = qn - take;
                    const unsigned short carry = (lane < rem) ? pq[32 + lane] : (unsigned short)0;
                    __syncwarp();
                    if (lane < rem) pq[lane] = carry;
                    qn = rem;
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();

    // ---- write-out: face_index rows (coalesced), coverage words
    for (int i = threadIdx.x; i < TILE * TILE; i += NTHREADS) {
        const int yl = i / TILE, xl = i % TILE;
        const unsigned lo = (unsigned)(keys[i] & 0xffffffffull);
        const bool cov = lo != 0xffffffffu;
        face_index[((long)b * is + (ty0 + yl)) * is + tx0 + xl] = cov ? (int)lo : -1;
        const unsigned m = __ballot_sync(0xffffffffu, cov);
        if (lane == 0) {
            roww[yl][xl >> 5] = m;
            if (cov_row) cov_row[((long)b * is + (ty0 + yl)) * (is / 32) + ((tx0 + xl) >> 5)] = m;
        }
    }
    __syncthreads();
    if (aa) {
        const int R = is / 2;
        const int rtop = R - 1 - (ty0 >> 1);
        for (int i = threadIdx.x; i < (TILE / 2) * (TILE / 2); i += NTHREADS) {
            const int m = i / (TILE / 2), cc = i % (TILE / 2);
            const int xl = 2 * cc;
            const unsigned a = roww[2 * m][xl >> 5] >> (xl & 31), c = roww[2 * m + 1][xl >> 5] >> (xl & 31);
            const float s = (float)((a & 1u) + ((a >> 1) & 1u) + (c & 1u) + ((c >> 1) & 1u));
            alpha[((long)b * R + (rtop - m)) * R + (tx0 >> 1) + cc] = s * 0.25f;
        }
    } else {
        for (int i = threadIdx.x; i < TILE * TILE; i += NTHREADS) {
            const int yl = i / TILE, xl = i % TILE;
            const unsigned a = roww[yl][xl >> 5] >> (xl & 31);
            alpha[((long)b * is + (is - 1 - ty0 - yl)) * is + tx0 + xl] = (a & 1u) ? 1.f : 0.f;
        }
    }
    if (cov_col && threadIdx.x < 2 * TILE) {
        const int xl = threadIdx.x % TILE, yw = threadIdx.x / TILE;
        unsigned word = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) word |= ((roww[yw * 32 + j][xl >> 5] >> (xl & 31)) & 1u) << j;
        cov_col[((long)b * is + (tx0 + xl)) * (is / 32) + (ty0 >> 5) + yw] = word;
    }
}

// ------------------------------------------------------------------------------------------ sweep masks
// m_row[b][0] = uncovered & grad < 0 ("missing coverage"), m_row[b][1] = covered & grad > 0 ("excess"),
// at raster resolution in the raster frame; m_col the same, transposed (one line per column).
__global__ void __launch_bounds__(NTHREADS)
grad_prep_kernel(const float *__restrict__ grad_alpha, const uint32_t *__restrict__ cov_row,
                 const uint32_t *__restrict__ cov_col, int is, int aa, uint32_t *__restrict__ m_row,
                 uint32_t *__restrict__ m_col) {
    __shared__ float gs[TILE][TILE + 1];
    __shared__ uint32_t a_row[TILE][2], a_col[TILE][2], out_w[4][TILE][2];
    const int b = blockIdx.y;
    const int tiles_x = is / TILE;
    const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = is / 32;
    {   // coverage words of the tile: 64 rows x 2 words, 64 columns x 2 words (one load per thread)
        const int l = (threadIdx.x >> 1) & 63, w = threadIdx.x & 1;
        if (threadIdx.x < 128) a_row[l][w] = cov_row[((long)b * is + (ty0 + l)) * W + (tx0 >> 5) + w];
        else a_col[l][w] = cov_col[((long)b * is + (tx0 + l)) * W + (ty0 >> 5) + w];
    }
    if (aa) {
        const int R = is / 2, rtop = R - 1 - (ty0 >> 1);
        for (int i = threadIdx.x; i < (TILE / 2) * (TILE / 2); i += NTHREADS) {
            const int m = i / (TILE / 2), cc = i % (TILE / 2);
            gs[m][cc] = grad_alpha[((long)b * R + (rtop - m)) * R + (tx0 >> 1) + cc];
        }
    } else {
        for (int i = threadIdx.x; i < TILE * TILE; i += NTHREADS) {
            const int yl = i / TILE, xl = i % TILE;
            gs[yl][xl] = grad_alpha[((long)b * is + (is - 1 - ty0 - yl)) * is + tx0 + xl];
        }
    }
    __syncthreads();
    const int sh = aa ? 1 : 0;
    const long plane = (long)is * W;
    for (int task = warp; task < 2 * TILE; task += NWARPS) {
        {   // row words: line = raster row, bit = x
            const int yl = task >> 1, w = task & 1, xl = w * 32 + lane;
            const float g = gs[yl >> sh][xl >> sh];
            const unsigned neg = __ballot_sync(0xffffffffu, g < 0.f), pos = __ballot_sync(0xffffffffu, g > 0.f);
            if (lane == 0) {
                const unsigned A = a_row[yl][w];
                out_w[0][yl][w] = neg & ~A;
                out_w[1][yl][w] = pos & A;
            }
        }
        {   // column words: line = raster column, bit = y
            const int xl = task >> 1, w = task & 1, yl = w * 32 + lane;
            const float g = gs[yl >> sh][xl >> sh];
            const unsigned neg = __ballot_sync(0xffffffffu, g < 0.f), pos = __ballot_sync(0xffffffffu, g > 0.f);
            if (lane == 0) {
                const unsigned A = a_col[xl][w];
                out_w[2][xl][w] = neg & ~A;
                out_w[3][xl][w] = pos & A;
            }
        }
    }
    __syncthreads();
    {   // one store per thread and plane
        const int l = (threadIdx.x >> 1) & 63, w = threadIdx.x & 1;
        if (threadIdx.x < 128) {
            const long o = ((long)b * 2 * is + (ty0 + l)) * W + (tx0 >> 5) + w;
            m_row[o] = out_w[0][l][w];
            m_row[o + plane] = out_w[1][l][w];
        } else {
            const long o = ((long)b * 2 * is + (tx0 + l)) * W + (ty0 >> 5) + w;
            m_col[o] = out_w[2][l][w];
            m_col[o + plane] = out_w[3][l][w];
        }
    }
}

// ------------------------------------------------------------------------------------------ sweep runs
// Run-length form of the four sweep masks: maximal stretches of consecutive set bits of one line that
// carry the same |grad| (the silhouette loss gradient is piecewise constant: 2 * norm * (rend - ref) takes a
// handful of values and is constant over whole missing / excess regions). A sweep then costs O(runs on the
// line) instead of O(set pixels). Lines with more than HM_RASTER_RUN_CAP runs keep the bit-line path.
//   runs       [B][4][is][HM_RASTER_RUN_CAP] uint2 {start | end << 16, |grad| bits}
//   run_counts [B][4][is]  (0xffffffff = overflow);  list 0 mn_row, 1 mp_row, 2 mn_col, 3 mp_col
constexpr int RCAP = HM_RASTER_RUN_CAP;
constexpr unsigned RUN_OVERFLOW = 0xffffffffu;

struct BwdCtx {
    const float *grad;  // grad_alpha of this image [R,R]
    int is, aa, R;
    float eps;
};

__device__ __forceinline__ float fetch_grad(const BwdCtx &c, int y, int x) {
    if (c.aa) return 0.25f * __ldg(c.grad + (long)((c.is - 1 - y) >> 1) * c.R + (x >> 1));
    return __ldg(c.grad + (long)(c.is - 1 - y) * c.R + x);
}

__global__ void __launch_bounds__(NTHREADS)
build_runs_kernel(const float *__restrict__ grad_alpha, const uint32_t *__restrict__ m_row,
                  const uint32_t *__restrict__ m_col, int is, int aa, uint2 *__restrict__ runs,
                  uint32_t *__restrict__ run_counts) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * NTHREADS + threadIdx.x;
    if (idx >= 4 * is) return;
    const int l4 = idx / is, line = idx % is, plane = l4 & 1, col = l4 >> 1;
    const int W = is / 32;
    const uint32_t *words = (col ? m_col : m_row) + (((long)b * 2 + plane) * is + line) * W;
    BwdCtx ctx;
    ctx.is = is; ctx.aa = aa; ctx.R = aa ? is / 2 : is; ctx.eps = 0.f;
    ctx.grad = grad_alpha + (long)b * ctx.R * ctx.R;
    uint2 *out = runs + (((long)b * 4 + l4) * is + line) * RCAP;
    int n = 0, rs = -1, re = -1;
    float rg = 0.f;
    for (int w = 0; w < W; ++w) {
        unsigned bits = words[w];
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            const int d1 = w * 32 + bit;
            const float g = fabsf(col ? fetch_grad(ctx, d1, line) : fetch_grad(ctx, line, d1));
            if (rs >= 0 && d1 == re + 1 && g == rg) {
                re = d1;
            } else {
                if (rs >= 0) {
                    if (n < RCAP) out[n] = make_uint2((unsigned)rs | ((unsigned)re << 16), __float_as_uint(rg));
                    ++n;
                }
                rs = re = d1;
                rg = g;
            }
        }
    }
    if (rs >= 0) {
        if (n < RCAP) out[n] = make_uint2((unsigned)rs | ((unsigned)re << 16), __float_as_uint(rg));
        ++n;
    }
    run_counts[((long)b * 4 + l4) * is + line] = n > RCAP ? RUN_OVERFLOW : (unsigned)n;
}

// ------------------------------------------------------------------------------------------ backward
// Visits the set bits of `line` (one raster row or column, W <= 32 words, `nz` = mask of its non-zero words)
// in [a, c] and accumulates the two vertex contributions of backward_pixel_map for the crossing (d0, d1_cross).
__device__ __forceinline__ void sweep(const uint32_t *line, unsigned nz, int a, int c, int axis, int d0,
                                      float d1_cross, float ka, float p0d0, float p1d0, const BwdCtx &ctx,
                                      float &acc0, float &acc1) {
    if (a > c) return;
    const int wa = a >> 5, wc = c >> 5;
    unsigned wm = nz & (0xffffffffu << wa) & (0xffffffffu >> (31 - wc));
    if (!wm) return;
    const float fd0 = (float)d0;
    const bool has0 = p1d0 != fd0, has1 = p0d0 != fd0;
    const float c0 = ka / (p1d0 - fd0), c1 = ka / (fd0 - p0d0);
    while (wm) {
        const int w = __ffs(wm) - 1;
        wm &= wm - 1;
        unsigned bits = line[w];
        if (w == wa) bits &= 0xffffffffu << (a & 31);
        if (w == wc) bits &= 0xffffffffu >> (31 - (c & 31));
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            const int d1 = w * 32 + bit;
            const float g = axis == 0 ? fetch_grad(ctx, d1, d0) : fetch_grad(ctx, d0, d1);
            const float diff = fabsf(g);
            const float dd = (float)d1 - d1_cross;
            if (has0) {
                float dist = c0 * dd * 2.f / (float)ctx.is;
                dist = (0.f < dist) ? dist + ctx.eps : dist - ctx.eps;
                acc0 -= diff / dist;
            }
            if (has1) {
                float dist = c1 * dd * 2.f / (float)ctx.is;
                dist = (0.f < dist) ? dist + ctx.eps : dist - ctx.eps;
                acc1 -= diff / dist;
            }
        }
    }
}

// psi(z2) - psi(z1) = sum_{k=0}^{n-1} 1 / (z1 + k) for z2 = z1 + n, z1 >= 8 (asymptotic series, error < 3e-10)
__device__ __forceinline__ float harmonic_span(float z1, float n) {
    const float z2 = z1 + n;
    const float i1 = __frcp_rn(z1), i2 = __frcp_rn(z2);
    const float a1 = i1 * i1, a2 = i2 * i2;
    float r = log1pf(n * i1);
    r += 0.5f * (i1 - i2);
    r += (1.f / 12.f) * (a1 - a2);
    r -= (1.f / 120.f) * (a1 * a1 - a2 * a2);
    r += (1.f / 252.f) * (a1 * a1 * a1 - a2 * a2 * a2);
    return r;
}

// One (crossing, run) work item: the part [s, e] of one run of one sweep line seen from the crossing at
// x = d1_cross of scan-line d0. Pixels within NEAR_PX of the crossing (the large terms) are evaluated one by
// one; beyond, every pixel of the run has the same weight G and the sum of 1 / dist is the harmonic sum
// G / K * sum 1 / (|d1 - x| + eps / |K|), taken in closed form (dist = K (d1 - x) +- eps, K = c * 2 / is).
constexpr float NEAR_PX = 8.f;
__device__ __forceinline__ void eval_item(float x, float c0, float c1, float G, int s, int e, bool has0, bool has1,
                                          float inv_is2, float eps, float &a0, float &a1) {
    const float K0 = c0 * inv_is2, K1 = c1 * inv_is2;
    const float rK0 = __frcp_rn(K0), rK1 = __frcp_rn(K1);
    const float del0 = eps * fabsf(rK0), del1 = eps * fabsf(rK1);
    const int near_lo = __float2int_rz(fmaxf(ceilf(x - NEAR_PX), -1.f));
    const int near_hi = __float2int_rz(fminf(floorf(x + NEAR_PX), 65535.f));
    float h0 = 0.f, h1 = 0.f;
    const int ns = max(s, near_lo), ne = min(e, near_hi);
    for (int d1 = ns; d1 <= ne; ++d1) {
        const float dd = (float)d1 - x;
        float dist0 = K0 * dd, dist1 = K1 * dd;
        dist0 = (0.f < dist0) ? dist0 + eps : dist0 - eps;
        dist1 = (0.f < dist1) ? dist1 + eps : dist1 - eps;
        h0 += __frcp_rn(dist0);
        h1 += __frcp_rn(dist1);
    }
    const int ps = max(s, near_hi + 1);  // far zone beyond the crossing: dist = K (dd + eps / |K|)
    if (ps <= e) {
        const float n = (float)(e - ps + 1), z = (float)ps - x;
        h0 += harmonic_span(z + del0, n) * rK0;
        h1 += harmonic_span(z + del1, n) * rK1;
    }
    const int me = min(e, near_lo - 1);  // far zone before the crossing: dist = -K (|dd| + eps / |K|)
    if (s <= me) {
        const float n = (float)(me - s + 1), z = x - (float)me;
        h0 -= harmonic_span(z + del0, n) * rK0;
        h1 -= harmonic_span(z + del1, n) * rK1;
    }
    a0 = has0 ? -G * h0 : 0.f;
    a1 = has1 ? -G * h1 : 0.f;
}

constexpr int WQCAP = 96;   // per-warp queue of (crossing, run) items
constexpr int BWD_LISTCAP = 1024;
struct WarpQueue {
    float x[WQCAP], G[WQCAP];
    unsigned se[WQCAP], meta[WQCAP];  // s | e << 16 ; owner task (0..255) | d0 << 8
};

// Parameters of the (up to) 256 tasks of a chunk, one per thread, readable by every thread of the CTA (the
// crossings of a task are evaluated by whichever lanes the load-balanced search hands them to).
struct TaskTable {
    float p0d0[NTHREADS], p0d1[NTHREADS], p1d0[NTHREADS], p2d0[NTHREADS], p2d1[NTHREADS], slope[NTHREADS],
        slope02[NTHREADS], slope21[NTHREADS], ka[NTHREADS];
    int packed[NTHREADS], fn[NTHREADS], incl[NTHREADS];
};

__device__ __forceinline__ void eval_queued(const TaskTable &tt, float x, float G, unsigned se, unsigned meta,
                                            float inv_is2, float eps, float &a0, float &a1) {
    const int o = meta & 255u;
    const float fd0 = (float)(meta >> 8);
    const float p0d0 = tt.p0d0[o], p1d0 = tt.p1d0[o], ka = tt.ka[o];
    eval_item(x, ka / (p1d0 - fd0), ka / (fd0 - p0d0), G, (int)(se & 0xffffu), (int)(se >> 16), p1d0 != fd0,
              p0d0 != fd0, inv_is2, eps, a0, a1);
}

// Per-task accumulators live in shared memory (float; atomics there are CAS loops, so contention must be
// avoided): a drain pass evaluates 32 items, sums the results of the items that belong to the same task inside
// the warp (match_any + shuffles), and one lane per task does a plain read-modify-write.
__device__ __forceinline__ void acc_add(float *slot, float v) {
    if (v != 0.f) atomicAdd(slot, v);
}

__device__ __forceinline__ void drain_queue(WarpQueue &q, const TaskTable &tt, int n, float (*wacc)[2], float inv_is2,
                                            float eps) {
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        float a0 = 0.f, a1 = 0.f;
        unsigned key = 256u + (unsigned)lane;  // lanes without an item form singleton groups
        if (i < n) {
            const unsigned meta = q.meta[i];
            eval_queued(tt, q.x[i], q.G[i], q.se[i], meta, inv_is2, eps, a0, a1);
            key = meta & 255u;
        }
        // segmented sum over runs of equal keys (items of one task sit next to each other in the queue almost
        // always); the head of every run adds to the task's accumulator, so the CAS loop of a shared-memory
        // float add rarely finds contention
        const unsigned knext = __shfl_down_sync(FULL, key, 1);
        const unsigned bnd = __ballot_sync(FULL, lane == 31 || knext != key);  // last lane of every run
        const int end = lane + __ffs(bnd >> lane) - 1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float v0 = __shfl_down_sync(FULL, a0, d), v1 = __shfl_down_sync(FULL, a1, d);
            if (lane + d <= end) { a0 += v0; a1 += v1; }
        }
        const unsigned kprev = __shfl_up_sync(FULL, key, 1);
        if ((lane == 0 || kprev != key) && key < 256u) {
            acc_add(&wacc[key][0], a0);
            acc_add(&wacc[key][1], a1);
        }
        __syncwarp();
    }
}

// One CTA per (image, 64x64 tile). The (face, edge, axis) tasks of the faces touching the tile are taken 256
// at a time, one per thread; their scan-line crossings are flattened over ALL lanes of the CTA (load-balanced
// search over the CTA-wide prefix sum of the task lengths, task parameters read from a shared table), and the
// (crossing, run) pairs that have something to sweep go through a per-warp queue so that the expensive part
// runs on dense warps.
__global__ void __launch_bounds__(NTHREADS)
raster_bwd_kernel(const FaceRec *__restrict__ recs, const FaceBox *__restrict__ boxes, int F, int V, int is, int aa,
                  float eps, const int32_t *__restrict__ face_index, const float *__restrict__ grad_alpha,
                  const uint32_t *__restrict__ cov_row, const uint32_t *__restrict__ cov_col,
                  const uint32_t *__restrict__ m_row, const uint32_t *__restrict__ m_col,
                  const uint2 *__restrict__ runs, const uint32_t *__restrict__ run_counts,
                  float *__restrict__ grad_ndc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int list[BWD_LISTCAP];
    __shared__ int cnt, next;
    __shared__ __align__(16) uint2 srun[4][TILE][RCAP];  // run lists: mn_row, mp_row, mn_col, mp_col
    __shared__ __align__(16) unsigned scount[4][TILE];
    __shared__ int wqn[NWARPS], wsum[NWARPS], wend[NWARPS];
    __shared__ float wacc[NTHREADS][2];
    __shared__ __align__(8) uint64_t bar;
    const int W = is / 32;
    const int LW = TILE * W;  // words per coverage block: a_row, a_col
    // dynamic shared memory: face_index tile | coverage lines | per-warp queues | task table of the chunk
    int *fi = reinterpret_cast<int *>(smem_raw);
    uint32_t *lines = reinterpret_cast<uint32_t *>(smem_raw + TILE * TILE * sizeof(int));
    WarpQueue *wq = reinterpret_cast<WarpQueue *>(smem_raw + TILE * TILE * sizeof(int) + 2 * LW * sizeof(uint32_t));
    TaskTable &tt = *reinterpret_cast<TaskTable *>(wq + NWARPS);
    const int b = blockIdx.y;
    const int tiles_x = is / TILE;
    const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    recs += (long)b * F;
    boxes += (long)b * F;
    grad_ndc += (long)b * V * 3;
    BwdCtx ctx;
    ctx.is = is; ctx.aa = aa; ctx.R = aa ? is / 2 : is; ctx.eps = eps;
    ctx.grad = grad_alpha + (long)b * ctx.R * ctx.R;
    const float inv_is2 = 2.f / (float)is;
    const long plane = (long)is * W;
    WarpQueue &q = wq[warp];

    int base = 0;
    bool staged = false;
    while (base < F) {
        const int n = next_batch(boxes, base, F, tx0, ty0, list, &cnt, &next, BWD_LISTCAP);
        if (n == 0) continue;
        if (!staged) {
            // ---- stage the tile's face_index rows, coverage lines and run lists with TMA bulk copies
            //      (tiles no face touches never get here: their face_index is never read)
            if (threadIdx.x == 0) {
                mbar_init(&bar, 1);
                mbar_fence_init();
            }
            __syncthreads();
            if (warp == 0) {
                if (lane == 0)
                    mbar_expect_tx(&bar, (uint32_t)(2 * LW * 4 + 4 * TILE * RCAP * 8 + 4 * TILE * 4));
                __syncwarp();
                if (lane == 0) tma_bulk_g2s(lines, cov_row + ((long)b * is + ty0) * W, (uint32_t)(LW * 4), &bar);
                if (lane == 1) tma_bulk_g2s(lines + LW, cov_col + ((long)b * is + tx0) * W, (uint32_t)(LW * 4), &bar);
                if (lane >= 8 && lane < 12) {
                    const int l4 = lane - 8, t0 = (l4 >> 1) ? tx0 : ty0;
                    tma_bulk_g2s(&srun[l4][0][0], runs + (((long)b * 4 + l4) * is + t0) * RCAP, TILE * RCAP * 8, &bar);
                    tma_bulk_g2s(&scount[l4][0], run_counts + ((long)b * 4 + l4) * is + t0, TILE * 4, &bar);
                }
            }
            // the face_index tile is 64 rows of 256 B: coalesced 16-byte loads (one bulk copy per row costs more
            // in TMA issue overhead than the bytes are worth)
            for (int i = threadIdx.x; i < TILE * TILE / 4; i += NTHREADS) {
                const int r = i / (TILE / 4), c4 = i % (TILE / 4);
                const int4 v = __ldg(reinterpret_cast<const int4 *>(face_index + ((long)b * is + ty0 + r) * is + tx0) + c4);
                reinterpret_cast<int4 *>(fi)[i] = v;
            }
            if (warp == 0) mbar_wait(&bar, 0);
            __syncthreads();
            staged = true;
        }
        const int ntasks = 6 * n;
        for (int chunk = 0; chunk < ntasks; chunk += NTHREADS) {
            // ---- one (face, edge, axis) task per thread; its parameters go to the CTA-wide table
            const int task = chunk + threadIdx.x;
            int vid0 = 0, vid1 = 0, axis = 0, len = 0;
            {
                float p0d0 = 0.f, p0d1 = 0.f, p1d0 = 0.f, p2d0 = 0.f, p2d1 = 0.f, slope = 0.f, slope02 = 0.f,
                      slope21 = 0.f, ka = 0.f;
                int fn = -1, dir = 1, lo = 0;
                if (task < ntasks) {
                    const int e = (task % 6) >> 1;
                    axis = task & 1;
                    const FaceRec *rp = recs + list[task / 6];
                    const float4 q0 = __ldg(reinterpret_cast<const float4 *>(rp));
                    const float4 q1 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
                    const float4 q2 = __ldg(reinterpret_cast<const float4 *>(rp) + 2);
                    const int4 q3 = __ldg(reinterpret_cast<const int4 *>(rp) + 3);
                    fn = __float_as_int(q2.y);
                    const int v0 = __float_as_int(q2.z), v1 = __float_as_int(q2.w), v2 = q3.x;
                    // pixel-space corners along (d0, d1) = (x, y) for axis 0, (y, x) for axis 1
                    const float ax = to_pix(q0.x, is), ay = to_pix(q0.y, is), bx = to_pix(q0.w, is), by = to_pix(q1.x, is),
                                cx = to_pix(q1.z, is), cy = to_pix(q1.w, is);
                    const float a0 = axis ? ay : ax, a1 = axis ? ax : ay, b0 = axis ? by : bx, b1 = axis ? bx : by,
                                c0 = axis ? cy : cx, c1 = axis ? cx : cy;
                    // vertex order of this edge: p0 = corner e, p1 = corner e+1, p2 = corner e+2
                    p0d0 = e == 0 ? a0 : e == 1 ? b0 : c0; p0d1 = e == 0 ? a1 : e == 1 ? b1 : c1;
                    p1d0 = e == 0 ? b0 : e == 1 ? c0 : a0;
                    const float p1d1 = e == 0 ? b1 : e == 1 ? c1 : a1;
                    p2d0 = e == 0 ? c0 : e == 1 ? a0 : b0; p2d1 = e == 0 ? c1 : e == 1 ? a1 : b1;
                    vid0 = e == 0 ? v0 : e == 1 ? v1 : v2; vid1 = e == 0 ? v1 : e == 1 ? v2 : v0;
                    if (axis == 0) dir = (p0d0 < p1d0) ? -1 : 1;
                    else dir = (p0d0 < p1d0) ? 1 : -1;
                    const int d0_from = __float2int_rz(fmaxf(ceilf(fminf(p0d0, p1d0)), 0.f));
                    const int d0_to = __float2int_rz(fminf(fmaxf(p0d0, p1d0), (float)(is - 1)));
                    const int t0 = axis == 0 ? tx0 : ty0;
                    lo = max(d0_from, t0);
                    len = max(min(d0_to, t0 + TILE - 1) - lo + 1, 0);
                    ka = p1d0 - p0d0;
                    slope = (p1d1 - p0d1) / ka;
                    slope02 = (p2d1 - p0d1) / (p2d0 - p0d0);
                    slope21 = (p1d1 - p2d1) / (p1d0 - p2d0);
                }
                // inclusive prefix sum of the task lengths over the CTA (warp scans + warp totals)
                int incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                __syncthreads();  // previous chunk fully consumed (table, accumulators, wsum)
                if (lane == 31) wsum[warp] = incl;
                const int t = threadIdx.x;
                tt.p0d0[t] = p0d0; tt.p0d1[t] = p0d1; tt.p1d0[t] = p1d0; tt.p2d0[t] = p2d0; tt.p2d1[t] = p2d1;
                tt.slope[t] = slope; tt.slope02[t] = slope02; tt.slope21[t] = slope21; tt.ka[t] = ka;
                tt.packed[t] = (lo & 0xffff) | (axis << 16) | ((dir > 0 ? 1 : 0) << 17);
                tt.fn[t] = fn;
                wacc[t][0] = 0.f;
                wacc[t][1] = 0.f;
                if (lane == 0) wqn[warp] = 0;
                __syncthreads();
                int offset = 0;
#pragma unroll
                for (int w = 0; w < NWARPS; ++w) offset += (w < warp) ? wsum[w] : 0;
                tt.incl[t] = incl + offset;
                if (lane == 31) wend[warp] = incl + offset;  // inclusive prefix at the end of each warp block
                __syncthreads();
            }
            const int total = tt.incl[NTHREADS - 1];

            // ---- crossings of the whole chunk, 32 per warp per round, dealt round-robin to the warps
            for (int kb = warp * 32; kb < total; kb += NTHREADS) {
                const int kk = kb + lane;
                const bool active = kk < total;
                // owner task of crossing kk = number of tasks whose inclusive prefix is <= kk: first the warp
                // block (8 independent compares against the block ends), then 5 dependent steps inside it
                int j = 0;
#pragma unroll
                for (int w = 0; w < NWARPS - 1; ++w) j += (wend[w] <= kk) ? 32 : 0;
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1)
                    if (tt.incl[j + sft - 1] <= kk) j += sft;
                const int o = active ? min(j, NTHREADS - 1) : NTHREADS - 1;
                const int o_excl = o > 0 ? tt.incl[o - 1] : 0, o_pack = tt.packed[o], o_fn = tt.fn[o];
                const float o_p0d0 = tt.p0d0[o], o_p0d1 = tt.p0d1[o], o_p1d0 = tt.p1d0[o], o_p2d0 = tt.p2d0[o];
                const float o_p2d1 = tt.p2d1[o], o_slope = tt.slope[o], o_s02 = tt.slope02[o], o_s21 = tt.slope21[o];
                const float o_ka = tt.ka[o];
                if (active) {
                    const int o_axis = (o_pack >> 16) & 1, o_dir = ((o_pack >> 17) & 1) ? 1 : -1;
                    const int d0 = (o_pack & 0xffff) + (kk - o_excl);
                    const int t0 = o_axis == 0 ? tx0 : ty0, t1 = o_axis == 0 ? ty0 : tx0;
                    const int l0 = d0 - t0;
                    const int lN = o_axis == 0 ? 2 : 0, lP = lN + 1;
                    const unsigned cN = scount[lN][l0], cP = scount[lP][l0];
                    if ((cN | cP) != 0u) {  // something to sweep on this line
                        const float fd0 = (float)d0;
                        const float x = o_slope * (fd0 - o_p0d0) + o_p0d1;
                        const int d1_in = __float2int_rz(o_dir > 0 ? floorf(x) : ceilf(x));
                        const int d1_out = d1_in + o_dir;
                        if (d1_in >= 0 && d1_in < is && d1_out >= 0 && d1_out < is && d1_in >= t1 && d1_in < t1 + TILE) {
                            const unsigned meta = (unsigned)o | ((unsigned)d0 << 8);
                            const int fs0 = o_axis == 0 ? 1 : TILE, fs1 = o_axis == 0 ? TILE : 1;
                            // sweep 0: out-sweep (missing-coverage list, from the out pixel to the border) when this
                            // face owns the in pixel; sweep 1: in-sweep (from the in pixel to the opposite edge)
#pragma unroll 1
                            for (int sw = 0; sw < 2; ++sw) {
                                int ls, ra, rc;
                                if (sw == 0) {
                                    if (cN == 0u || fi[l0 * fs0 + (d1_in - t1) * fs1] != o_fn) continue;
                                    const int lim = o_dir > 0 ? is - 1 : 0;
                                    ls = lN; ra = min(d1_out, lim); rc = max(d1_out, lim);
                                } else {
                                    const uint32_t *A = lines + (o_axis == 0 ? LW : 0) + l0 * W;
                                    const bool alpha_out = (A[d1_out >> 5] >> (d1_out & 31)) & 1u;
                                    ls = alpha_out ? lN : lP;
                                    if ((alpha_out ? cN : cP) == 0u) continue;
                                    float c2;
                                    if ((fd0 - o_p0d0) * (fd0 - o_p2d0) < 0.f) c2 = o_s02 * (fd0 - o_p0d0) + o_p0d1;
                                    else c2 = o_s21 * (fd0 - o_p2d0) + o_p2d1;
                                    const int lim = __float2int_rz(o_dir > 0 ? ceilf(c2) : floorf(c2));
                                    ra = max(min(d1_in, lim), 0); rc = min(max(d1_in, lim), is - 1);
                                }
                                const unsigned cs = ls == lN ? cN : cP;
                                if (ra > rc) continue;
                                if (cs == RUN_OVERFLOW) {
                                    // more runs than the list holds: walk the bit line (global memory)
                                    const uint32_t *line = ((ls >> 1) ? m_col : m_row) +
                                                           ((long)b * 2 * is + (ls >> 1 ? tx0 : ty0) + l0) * W + (ls & 1) * plane;
                                    float a0 = 0.f, a1 = 0.f;
                                    sweep(line, FULL, ra, rc, o_axis, d0, x, o_ka, o_p0d0, o_p1d0, ctx, a0, a1);
                                    acc_add(&wacc[o][0], a0);
                                    acc_add(&wacc[o][1], a1);
                                    continue;
                                }
                                for (unsigned r = 0; r < cs; ++r) {
                                    const uint2 run = srun[ls][l0][r];
                                    const int s = max(ra, (int)(run.x & 0xffffu)), e = min(rc, (int)(run.x >> 16));
                                    if (s > e) continue;
                                    const unsigned se = (unsigned)s | ((unsigned)e << 16);
                                    const int pos = atomicAdd(&wqn[warp], 1);
                                    if (pos < WQCAP) {
                                        q.x[pos] = x; q.G[pos] = __uint_as_float(run.y); q.se[pos] = se; q.meta[pos] = meta;
                                    } else {  // queue full: evaluate in place
                                        float a0, a1;
                                        eval_queued(tt, x, __uint_as_float(run.y), se, meta, inv_is2, eps, a0, a1);
                                        acc_add(&wacc[o][0], a0);
                                        acc_add(&wacc[o][1], a1);
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                const int nq = min(wqn[warp], WQCAP);
                if (nq > WQCAP - 48 || kb + NTHREADS >= total) {
                    drain_queue(q, tt, nq, wacc, inv_is2, eps);
                    __syncwarp();
                    if (lane == 0) wqn[warp] = 0;
                    __syncwarp();
                }
            }
            __syncthreads();
            // slot pi0*3 + (1 - axis): axis 0 sweeps along y and yields the y gradient, axis 1 the x gradient
            const float acc0 = wacc[threadIdx.x][0], acc1 = wacc[threadIdx.x][1];
            if (acc0 != 0.f) atomicAdd(grad_ndc + (long)vid0 * 3 + (1 - axis), acc0);
            if (acc1 != 0.f) atomicAdd(grad_ndc + (long)vid1 * 3 + (1 - axis), acc1);
        }
    }
}

// ------------------------------------------------------------------------------------------ silhouette loss
__global__ void __launch_bounds__(NTHREADS)
sil_loss_kernel(const float *__restrict__ alpha, const int8_t *__restrict__ target, const float *__restrict__ norm,
                float weight, int npix, float *__restrict__ loss_img, int loss_stride, float *__restrict__ iou_img,
                int iou_stride, float *__restrict__ grad_alpha) {
    __shared__ float scratch[3 * 32];
    const int b = blockIdx.x;
    const float nb = norm[b];
    const float gscale = 2.f * weight * nb;
    const float4 *a4 = reinterpret_cast<const float4 *>(alpha + (long)b * npix);
    const char4 *t4 = reinterpret_cast<const char4 *>(target + (long)b * npix);
    float4 *g4 = reinterpret_cast<float4 *>(grad_alpha + (long)b * npix);
    float acc[3] = {0.f, 0.f, 0.f};  // sum sq, intersection, union
    for (int i = threadIdx.x; i < npix / 4; i += NTHREADS) {
        const float4 a = a4[i];
        const char4 t = t4[i];
        const float av[4] = {a.x, a.y, a.z, a.w};
        const int tv[4] = {t.x, t.y, t.z, t.w};
        float gv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float keep = tv[k] >= 0 ? 1.f : 0.f, ref = tv[k] > 0 ? 1.f : 0.f;
            const float img = keep * av[k];
            const float d = img - ref;
            acc[0] += d * d;
            acc[1] += img * ref;
            acc[2] += fminf(fmaxf(img + ref, 0.f), 1.f);
            gv[k] = gscale * keep * d;
        }
        if (grad_alpha) g4[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
    }
    block_sum<3>(acc, scratch);
    if (threadIdx.x == 0) {
        if (loss_img) loss_img[(long)b * loss_stride] = acc[0] * nb;
        if (iou_img) iou_img[(long)b * iou_stride] = acc[1] / (acc[2] + 1e-6f);
    }
}

int check_raster_size(int image_size, int aa, int *is_out) {
    const int is = aa ? 2 * image_size : image_size;
    HM_REQUIRE(image_size > 0, "image_size must be positive");
    HM_UNSUPPORTED(is % TILE != 0 || is > 16384, "raster size %d must be a multiple of %d (<= 16384)", is, TILE);
    *is_out = is;
    return HM_OK;
}

}  // namespace

extern "C" {

int hm_project_fwd(const float *verts, const float *K, int K_batch, const float *R, const float *t,
                   const float *dist, int dist_batch, float orig_size, float eps, int B, int V, float *ndc,
                   void *stream) {
    HM_REQUIRE(verts && K && ndc, "hm_project_fwd: null pointer");
    HM_REQUIRE(B >= 0 && V >= 0 && (K_batch == 1 || K_batch == B), "hm_project_fwd: bad sizes B=%d V=%d K_batch=%d", B, V, K_batch);
    if ((long)B * V == 0) return HM_OK;
    const long n = (long)B * V;
    project_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, hm_stream(stream)>>>(verts, K, K_batch, R, t, dist,
                                                                                  dist_batch, orig_size, eps, B, V, ndc);
    HM_CHECK_LAUNCH("hm_project_fwd");
    return HM_OK;
}

int hm_project_bwd(const float *verts, const float *K, int K_batch, const float *R, const float *t,
                   float orig_size, float eps, int B, int V, const float *grad_ndc, float *grad_verts,
                   int accumulate, void *stream) {
    HM_REQUIRE(verts && K && grad_ndc && grad_verts, "hm_project_bwd: null pointer");
    HM_REQUIRE(B >= 0 && V >= 0 && (K_batch == 1 || K_batch == B), "hm_project_bwd: bad sizes");
    if ((long)B * V == 0) return HM_OK;
    const long n = (long)B * V;
    project_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, hm_stream(stream)>>>(verts, K, K_batch, R, t, orig_size,
                                                                                  eps, B, V, grad_ndc, grad_verts,
                                                                                  accumulate);
    HM_CHECK_LAUNCH("hm_project_bwd");
    return HM_OK;
}

int hm_raster_setup(const float *ndc, const int32_t *faces, int faces_batch, int B, int V, int F,
                    int image_size, int anti_aliasing, int fill_back, void *records, void *bboxes,
                    void *stream) {
    HM_REQUIRE(ndc && faces && records && bboxes, "hm_raster_setup: null pointer");
    HM_REQUIRE(B >= 0 && V > 0 && F >= 0 && (faces_batch == 1 || faces_batch == B), "hm_raster_setup: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if ((long)B * F == 0) return HM_OK;
    const long n = (long)B * F;
    face_setup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, hm_stream(stream)>>>(
        ndc, faces, faces_batch, B, V, F, is, fill_back, static_cast<FaceRec *>(records),
        static_cast<FaceBox *>(bboxes));
    HM_CHECK_LAUNCH("hm_raster_setup");
    return HM_OK;
}

int hm_raster_sil_fwd(const void *records, const void *bboxes, int B, int F, int image_size,
                      int anti_aliasing, float near_, float far_, int32_t *face_index, float *alpha,
                      uint32_t *cov_row, uint32_t *cov_col, void *stream) {
    HM_REQUIRE(records && bboxes && face_index && alpha, "hm_raster_sil_fwd: null pointer");
    HM_REQUIRE(B >= 0 && F >= 0 && B <= 65535, "hm_raster_sil_fwd: bad sizes (B <= 65535)");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0) return HM_OK;
    dim3 grid((is / TILE) * (is / TILE), B);
    raster_fwd_kernel<<<grid, NTHREADS, 0, hm_stream(stream)>>>(
        static_cast<const FaceRec *>(records), static_cast<const FaceBox *>(bboxes), F, is, anti_aliasing, near_, far_,
        face_index, alpha, cov_row, cov_col);
    HM_CHECK_LAUNCH("hm_raster_sil_fwd");
    return HM_OK;
}

int hm_raster_grad_prep(const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, int B,
                        int image_size, int anti_aliasing, uint32_t *m_row, uint32_t *m_col, void *runs,
                        uint32_t *run_counts, void *stream) {
    HM_REQUIRE(grad_alpha && cov_row && cov_col && m_row && m_col && runs && run_counts,
               "hm_raster_grad_prep: null pointer");
    HM_REQUIRE(B >= 0 && B <= 65535, "hm_raster_grad_prep: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0) return HM_OK;
    dim3 grid((is / TILE) * (is / TILE), B);
    grad_prep_kernel<<<grid, NTHREADS, 0, hm_stream(stream)>>>(grad_alpha, cov_row, cov_col, is, anti_aliasing, m_row,
                                                               m_col);
    HM_CHECK_LAUNCH("hm_raster_grad_prep");
    dim3 grid2((4 * is + NTHREADS - 1) / NTHREADS, B);
    build_runs_kernel<<<grid2, NTHREADS, 0, hm_stream(stream)>>>(grad_alpha, m_row, m_col, is, anti_aliasing,
                                                                 static_cast<uint2 *>(runs), run_counts);
    HM_CHECK_LAUNCH("hm_raster_grad_prep(runs)");
    return HM_OK;
}

int hm_raster_sil_bwd(const void *records, const void *bboxes, const int32_t *face_index,
                      const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col,
                      const uint32_t *m_row, const uint32_t *m_col, const void *runs, const uint32_t *run_counts,
                      int B, int V, int F, int image_size, int anti_aliasing, float eps, float *grad_ndc,
                      void *stream) {
    HM_REQUIRE(records && bboxes && face_index && grad_alpha && cov_row && cov_col && m_row && m_col && runs &&
                   run_counts && grad_ndc,
               "hm_raster_sil_bwd: null pointer");
    HM_REQUIRE(B >= 0 && F >= 0 && V > 0 && B <= 65535, "hm_raster_sil_bwd: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0 || F == 0) return HM_OK;
    const size_t smem = (size_t)TILE * TILE * 4 + (size_t)2 * TILE * (is / 32) * 4 + NWARPS * sizeof(WarpQueue) +
                        sizeof(TaskTable);
    HM_UNSUPPORTED(is > 1024, "hm_raster_sil_bwd: raster size %d > 1024 is not supported", is);
    static size_t configured = 0;  // static + dynamic shared memory exceeds the 48 KB default: opt in once per size
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(raster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            hm_set_error("hm_raster_sil_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return HM_ERR_CUDA;
        }
        configured = smem;
    }
    dim3 grid((is / TILE) * (is / TILE), B);
    raster_bwd_kernel<<<grid, NTHREADS, smem, hm_stream(stream)>>>(
        static_cast<const FaceRec *>(records), static_cast<const FaceBox *>(bboxes), F, V, is, anti_aliasing, eps,
        face_index, grad_alpha, cov_row, cov_col, m_row, m_col, static_cast<const uint2 *>(runs), run_counts, grad_ndc);
    HM_CHECK_LAUNCH("hm_raster_sil_bwd");
    return HM_OK;
}

int hm_sil_loss_fwd_bwd(const float *alpha, const int8_t *target, const float *norm, float weight, int B,
                        int image_size, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                        float *grad_alpha, void *stream) {
    HM_REQUIRE(alpha && target && norm, "hm_sil_loss_fwd_bwd: null pointer");
    HM_REQUIRE(B >= 0 && image_size > 0 && (image_size * image_size) % 4 == 0, "hm_sil_loss_fwd_bwd: bad sizes");
    if (B == 0) return HM_OK;
    sil_loss_kernel<<<B, NTHREADS, 0, hm_stream(stream)>>>(alpha, target, norm, weight, image_size * image_size,
                                                           loss_img, loss_stride, iou_img, iou_stride, grad_alpha);
    HM_CHECK_LAUNCH("hm_sil_loss_fwd_bwd");
    return HM_OK;
}

}  // extern "C"
