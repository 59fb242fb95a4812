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

// Named knob, parity unpinned (the un-vendored package is not available): upstream's face set-up reportedly keeps
// the barycentric determinant away from zero (|det| >= 1e-10) before dividing by it. 0 = divide by the raw determinant.
// The oracle carries the same constant (NMR_DET_CLAMP in oracle/csrc/nmr_raster.c); they must be changed together.
#ifndef HM_RASTER_DET_CLAMP
#define HM_RASTER_DET_CLAMP 0.f
#endif
constexpr int TILE = 64;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

struct __align__(16) FaceRec {
    float c[9];   // x0 y0 z0 x1 y1 z1 x2 y2 z2 in NDC, front-facing winding
    int fn;       // face number in the doubled (fill_back) numbering, -1 = culled
    int v[3];     // vertex indices in that winding
    short bb[4];  // pixel bbox x0 y0 x1 y1 (clamped), empty when x0 > x1
    int pad;      // bit 0: both windings are front-facing (degenerate face): forward and backward also process F + f
    float inv[9];  // barycentric matrix in pixel coordinates, already divided by its determinant
    float iz[3];   // 1 / z of the three corners (fast depth of the forward pass)
    int exact;     // 1: some corner depth is not a plain positive float -> forward uses the reference arithmetic only
    short fb[4];   // forward box x0 y0 x1 y1: the rows / columns the forward visits (tight; = bb for slivers, see the setup)
    float pad2;    // min corner depth
};
static_assert(sizeof(FaceRec) == 128, "record size");

// Backward record: pixel-space corners of the stored winding and the slope of every edge along both sweep axes
// (the quotients backward_pixel_map evaluates per crossing, computed once with the same rounded operations).
struct __align__(16) BwdRec {
    float px[3], py[3];   // to_pix of the corners
    float slope[3][2];    // edge e = corner e -> e + 1: [0] = dy / dx (axis 0: lines x = d0), [1] = dx / dy (axis 1)
    int v[3];             // vertex indices
    int meta;             // < 0: culled / off screen; else fn | BWD_IRREGULAR | BWD_BOTH
    uint32_t span[3][2];  // per (edge, axis): scan-lines d0_from | d0_to << 12 (empty: 1, 0), sweep direction > 0 << 24,
                          //   steep (|slope| > SMAX or not finite) << 25
    uint32_t pad[2];
};
static_assert(sizeof(BwdRec) == 96 && sizeof(FaceRec) + sizeof(BwdRec) == HM_FACE_RECORD_BYTES, "record size");
constexpr int BWD_FN_MASK = (1 << 29) - 1;
constexpr int BWD_IRREGULAR = 1 << 29;   // a corner on an integer pixel coordinate, in (-1, 0), or not finite: never skipped
constexpr int BWD_BOTH = 1 << 30;        // both windings are front-facing: F + f is differentiated too
constexpr float BWD_SMAX = 128.f;        // tasks steeper than this are enumerated from the face (see the backward kernels)

struct __align__(8) FaceBox {
    short x0, y0, x1, y1;   // y1 also carries the forward's pass bit (FBOX_REV): readers mask it off
};
constexpr int FBOX_MASK = 0xfff;   // image sizes <= 4096 (the backward spans pack scan-lines in 12 bits too)
constexpr int FBOX_REV = 1 << 12;  // the stored winding is the reversed (fill_back) copy: second pass of the forward

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
                                  BwdRec *__restrict__ brecs, FaceBox *__restrict__ boxes, int *__restrict__ img_box) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < (long)B * F;   // (idle threads of the last block repeat the last face: same box, same stores)
    if (!live) i = (long)B * F - 1;
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
    bool rev = false, both = false;
    // back-face test of the original winding, then of the reversed copy appended by fill_back
    const bool front_orig = !((p[2][1] - p[0][1]) * (p[1][0] - p[0][0]) < (p[1][1] - p[0][1]) * (p[2][0] - p[0][0]));
    const bool front_rev = fill_back &&
                           !((p[0][1] - p[2][1]) * (p[1][0] - p[2][0]) < (p[1][1] - p[2][1]) * (p[0][0] - p[2][0]));
    if (front_orig) {
        fn = f;
        both = front_rev;  // signed area exactly 0 in fp32: the reference keeps BOTH copies (f and F + f)
    } else if (front_rev) {
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
    r.pad = both ? 1 : 0;
    int fwd_key = 0;
    {   // barycentric matrix of the stored winding (same expressions as the oracle's per-face setup)
        const float p00 = to_pix(r.c[0], is), p01 = to_pix(r.c[1], is), p10 = to_pix(r.c[3], is), p11 = to_pix(r.c[4], is),
                    p20 = to_pix(r.c[6], is), p21 = to_pix(r.c[7], is);
        const float m[9] = {p11 - p21, p20 - p10, p10 * p21 - p20 * p11,
                            p21 - p01, p00 - p20, p20 * p01 - p00 * p21,
                            p01 - p11, p10 - p00, p00 * p11 - p10 * p01};
        float den = p20 * (p01 - p11) + p00 * (p11 - p21) + p10 * (p21 - p01);
        if (HM_RASTER_DET_CLAMP > 0.f) {   // (see the constant: 0 = the raw determinant, as the oracle)
            if (den > 0.f) den = fmaxf(den, HM_RASTER_DET_CLAMP);
            else den = fminf(den, -HM_RASTER_DET_CLAMP);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) r.inv[k] = m[k] / den;
        r.exact = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float z = r.c[3 * k + 2];
            r.iz[k] = 1.f / z;
            if (!(z > 1e-20f && z < 1e20f)) r.exact = 1;
        }
        r.pad2 = fminf(fminf(r.c[2], r.c[5]), r.c[8]);   // nearest corner (hidden-layer test of the forward)
        // Forward box: a sample a whole pixel outside the bounding box of the corners cannot pass the three edge
        // predicates unless the triangle is a needle (the margin of the test is the sample's distance to the edge
        // lines, ~sin(apex angle) pixels, against ~1e-5 of rounding): faces with |det| >= 1e-3 |longest edge|^2 get the
        // tight box floor(min) .. ceil(max), needles and both-winding faces keep the reference's box with 1 px of slack.
        const float l2 = fmaxf(fmaxf((p10 - p00) * (p10 - p00) + (p11 - p01) * (p11 - p01),
                                     (p20 - p10) * (p20 - p10) + (p21 - p11) * (p21 - p11)),
                               (p00 - p20) * (p00 - p20) + (p01 - p21) * (p01 - p21));
        int g0 = x0, g1 = y0, g2 = x1, g3 = y1;
        if (x0 <= x1 && !both && fabsf(den) >= 1e-3f * l2 && l2 < 1e30f) {
            g0 = max(__float2int_rz(floorf(xmin)), 0); g1 = max(__float2int_rz(floorf(ymin)), 0);
            g2 = min(__float2int_rz(ceilf(xmax)), is - 1); g3 = min(__float2int_rz(ceilf(ymax)), is - 1);
        }
        r.fb[0] = (short)g0; r.fb[1] = (short)g1; r.fb[2] = (short)g2; r.fb[3] = (short)g3;
        fwd_key = rev ? FBOX_REV : 0;
    }
    recs[i] = r;
    {
        BwdRec br;
        bool irregular = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            br.px[k] = to_pix(r.c[3 * k], is);
            br.py[k] = to_pix(r.c[3 * k + 1], is);
            br.v[k] = r.v[k];
            irregular |= !(fabsf(br.px[k]) < 1e30f) || !(fabsf(br.py[k]) < 1e30f) || br.px[k] == floorf(br.px[k]) ||
                         br.py[k] == floorf(br.py[k]);
            // a coordinate in (-1, 0): backward_pixel_map truncates the end of a scan-line range towards zero, so an
            // edge that ends there still gets scan-line 0, outside the triangle (sweep ends extrapolated)
            irregular |= (br.px[k] > -1.f && br.px[k] < 0.f) || (br.py[k] > -1.f && br.py[k] < 0.f);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int e1 = e == 2 ? 0 : e + 1;
            br.slope[e][0] = (br.py[e1] - br.py[e]) / (br.px[e1] - br.px[e]);
            br.slope[e][1] = (br.px[e1] - br.px[e]) / (br.py[e1] - br.py[e]);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int e1 = e == 2 ? 0 : e + 1;
#pragma unroll
            for (int axis = 0; axis < 2; ++axis) {   // the scan-line range and direction backward_pixel_map derives
                const float p0 = axis ? br.py[e] : br.px[e], p1 = axis ? br.py[e1] : br.px[e1];
                int from = __float2int_rz(fmaxf(ceilf(fminf(p0, p1)), 0.f));
                int to = __float2int_rz(fminf(fmaxf(p0, p1), (float)(is - 1)));
                if (from > to) { from = 1; to = 0; }
                const bool plus = axis == 0 ? !(p0 < p1) : (p0 < p1);
                const bool steep = !(fabsf(br.slope[e][axis]) <= BWD_SMAX);
                br.span[e][axis] = (unsigned)from | ((unsigned)to << 12) | (plus ? 1u << 24 : 0u) | (steep ? 1u << 25 : 0u);
            }
        }
        br.pad[0] = br.pad[1] = 0u;
        br.meta = fn < 0 ? -1 : (fn | (irregular ? BWD_IRREGULAR : 0) | (both ? BWD_BOTH : 0));
        brecs[i] = br;
    }
    FaceBox bx;
    bx.x0 = (short)x0; bx.y0 = (short)y0; bx.x1 = (short)x1; bx.y1 = (short)(y1 | fwd_key);
    boxes[i] = bx;
    // pixel box of the whole image (all its faces): {x0, y0, -x1, -y1} under atomicMin, preset to 0x7f7f7f7f. One
    // reduction per warp over the lanes of the first lane's image, the other lanes (next image) on their own.
    {
        const unsigned FULL = 0xffffffffu;
        const bool has = x0 <= x1;
        const int b0 = __shfl_sync(FULL, b, 0);
        int v[4] = {has ? x0 : 0x7f7f7f7f, has ? y0 : 0x7f7f7f7f, has ? -x1 : 0x7f7f7f7f, has ? -y1 : 0x7f7f7f7f};
        const bool mine = b == b0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = __reduce_min_sync(FULL, mine ? v[k] : 0x7f7f7f7f);
            if ((threadIdx.x & 31) == 0) v[k] = m;
            if (((threadIdx.x & 31) == 0 || !mine) && v[k] != 0x7f7f7f7f && v[k] < img_box[4 * b + k]) atomicMin(&img_box[4 * b + k], v[k]);
        }
    }
}

// 1600 entries (with the 32 KB z-buffer: 44.4 KB per CTA) and 48 registers let five CTAs share an SM (40 warps; the
// kernel is latency-bound): 5 x (44.4 + 1) KB is just inside the 227 KB an SM gives its CTAs.
#ifndef HM_FWD_LISTCAP
#define HM_FWD_LISTCAP 1600
#endif
constexpr int LISTCAP = HM_FWD_LISTCAP;  // faces of one tile processed per batch

constexpr int SCAN = 4 * NTHREADS;  // faces tested per scan step (four 8-byte boxes per thread)
static_assert(LISTCAP >= SCAN, "a batch must hold one scan step (the scan continues while n <= LISTCAP - SCAN)");

// Appends the faces of [base, base + SCAN) whose bbox touches the tile: faces stored in their original winding (first
// pass of the forward) from the front of the list, reversed copies (second pass) from its back; cnt[0], cnt[1] count
// them.
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
    unsigned hits = 0, revs = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y1 = bx[k].y1 & FBOX_MASK;
        const bool hit = bx[k].x0 <= bx[k].x1 && bx[k].x0 <= tx0 + TILE - 1 && bx[k].x1 >= tx0 &&
                         bx[k].y0 <= ty0 + TILE - 1 && y1 >= ty0;
        hits |= (hit ? 1u : 0u) << k;
        revs |= ((hit && (bx[k].y1 & FBOX_REV)) ? 1u : 0u) << k;
    }
    const int mine = __popc(hits & ~revs) | (__popc(revs) << 16);   // both counts in one scan
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int wb0 = 0, wb1 = 0;
    if (lane == 31) {
        if (total & 0xffff) wb0 = atomicAdd(&cnt[0], total & 0xffff);
        if (total >> 16) wb1 = atomicAdd(&cnt[1], total >> 16);
    }
    wb0 = __shfl_sync(0xffffffffu, wb0, 31);
    wb1 = __shfl_sync(0xffffffffu, wb1, 31);
    int pos0 = wb0 + ((incl - mine) & 0xffff), pos1 = wb1 + ((incl - mine) >> 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((hits >> k) & 1u) {
            if ((revs >> k) & 1u) list[LISTCAP - 1 - pos1++] = f0 + k;
            else list[pos0++] = f0 + k;
        }
}

// Fills the list with the next batch of faces touching the tile (scanning from *base). Block-uniform; returns the
// number of entries (cnt[0] from the front + cnt[1] from the back).
__device__ __forceinline__ int next_batch(const FaceBox *__restrict__ boxes, int &base, int F, int tx0, int ty0,
                                          int *list, int *cnt, int *next) {
    __syncthreads();
    if (threadIdx.x == 0) { cnt[0] = 0; cnt[1] = 0; *next = 0; }
    __syncthreads();
    int n = 0;
    while (base < F && n <= LISTCAP - SCAN) {
        append_faces(boxes, base, F, tx0, ty0, list, cnt);
        base += SCAN;
        __syncthreads();
        n = cnt[0] + cnt[1];
    }
    return n;
}
// Entry li of the batch (front entries first).
__device__ __forceinline__ int batch_face(const int *list, int n0, int li) {
    return li < n0 ? list[li] : list[LISTCAP - 1 - (li - n0)];
}

// ------------------------------------------------------------------------------------------ forward
// Depth of the winner search. The reference evaluates, per covered sample, zp = 1 / (w0/ws/z0 + w1/ws/z1 + w2/ws/z2)
// with seven IEEE divisions. The kernel ranks samples with zp~ = ws / (w0*iz0 + w1*iz1 + w2*iz2) (one MUFU.RCP), which
// is within a few ulp of the reference value, and records as AMBIGUOUS every pixel at which two samples (or a sample
// and the near / far planes) came within AMB ulp of each other: only those pixels can rank differently under the
// reference arithmetic, and they are re-resolved from scratch with it after the main passes. face_index is
// therefore bit-exact while the common sample costs ~40 instructions instead of ~200.
constexpr unsigned AMB = 48;   // ulp; the two evaluation orders differ by < 8 ulp each
constexpr int HZ = 4;          // hierarchical-z block edge (pixels)
constexpr int HZN = TILE / HZ;
#ifndef HM_FWD_FILTER_MIN
#define HM_FWD_FILTER_MIN 32   // hidden-layer entries of a tile from which the thread-per-entry filter pays
#endif
constexpr int AMBCAP = NWARPS * 64;

__device__ __forceinline__ bool ulp_close(unsigned a, unsigned b) {
    const unsigned d = a > b ? a - b : b - a;
    return d <= AMB;
}

// The reference's per-sample arithmetic (oracle/csrc/nmr_raster.c), individually rounded operations.
__device__ __forceinline__ float exact_depth(const float *inv, int xi, int yi, float z0, float z1, float z2) {
    float w0 = inv[0] * xi + inv[1] * yi + inv[2];
    float w1 = inv[3] * xi + inv[4] * yi + inv[5];
    float w2 = inv[6] * xi + inv[7] * yi + inv[8];
    w0 = fminf(fmaxf(w0, 0.f), 1.f);
    w1 = fminf(fmaxf(w1, 0.f), 1.f);
    w2 = fminf(fmaxf(w2, 0.f), 1.f);
    const float ws = w0 + w1 + w2;
    w0 /= ws; w1 /= ws; w2 /= ws;
    return 1.f / (w0 / z0 + w1 / z1 + w2 / z2);
}

// Coverage of one raster row by one edge. The reference skips a sample when A < (xp - xk) * eky with
// A = (yp - yk) * ekx (each operation rounded). For a fixed row the right-hand side is a monotone function of the
// pixel index (rounding is monotone), so the samples that pass form a prefix (eky > 0), a suffix (eky < 0) or all /
// none (eky == 0) of the row: the boundary is located from the analytic crossing and pinned with the reference
// predicate itself, which makes the interval exact without testing every sample.
struct RowCtx {
    int is;
    bool pow2;
    float inv_is;
};
__device__ __forceinline__ bool edge_pass(const RowCtx &rc, float A, float xk, float eky, int xi) {
    const float xp = rc.pow2 ? (float)(2 * xi + 1 - rc.is) * rc.inv_is : (float)(2 * xi + 1 - rc.is) / (float)rc.is;
    return !(A < (xp - xk) * eky);
}
__device__ __forceinline__ void clip_edge(const RowCtx &rc, float A, float xk, float eky, int X0, int X1, int &lo,
                                          int &hi) {
    if (eky > 0.f) {
        const float est = to_pix(xk + __fdividef(A, eky), rc.is);
        int c = (est == est) ? __float2int_rz(fminf(fmaxf(floorf(est), (float)(X0 - 1)), (float)X1)) : X1;
        while (c < X1 && edge_pass(rc, A, xk, eky, c + 1)) ++c;
        while (c >= X0 && !edge_pass(rc, A, xk, eky, c)) --c;
        hi = min(hi, c);
    } else if (eky < 0.f) {
        const float est = to_pix(xk + __fdividef(A, eky), rc.is);
        int c = (est == est) ? __float2int_rz(fminf(fmaxf(ceilf(est), (float)X0), (float)(X1 + 1))) : X0;
        while (c > X0 && edge_pass(rc, A, xk, eky, c - 1)) --c;
        while (c <= X1 && !edge_pass(rc, A, xk, eky, c)) ++c;
        lo = max(lo, c);
    } else if (eky == 0.f) {
        if (A < 0.f) { lo = 1; hi = 0; }  // (xp - xk) * (+-0) = +-0
    }  // NaN slope: the comparison is false for every sample, all pass
}

// A tile no face touches: constants, four 16-byte stores per thread.
__device__ __forceinline__ void write_untouched_tile(int b, int tx0, int ty0, int is, int aa,
                                                     int32_t *__restrict__ face_index, float *__restrict__ alpha,
                                                     uint32_t *__restrict__ cov_row, uint32_t *__restrict__ cov_col,
                                                     unsigned char *__restrict__ cov_blocks) {
    const int W = is / 32, t = threadIdx.x;
    {
        int4 *p = reinterpret_cast<int4 *>(face_index + ((long)b * is + ty0 + (t >> 4)) * is + tx0 + 4 * (t & 15));
        const long rs = 4 * (long)is;   // 16 rows further, in int4 units
        const int4 m1 = make_int4(-1, -1, -1, -1);
        p[0] = m1; p[rs] = m1; p[2 * rs] = m1; p[3 * rs] = m1;
    }
    if (t < 2 * TILE) {
        const int yl = t >> 1, w = t & 1;
        if (cov_row) cov_row[((long)b * is + (ty0 + yl)) * W + (tx0 >> 5) + w] = 0u;
        if (cov_col) cov_col[((long)b * is + (tx0 + yl)) * W + (ty0 >> 5) + w] = 0u;
    }
    if (cov_blocks && t < 2 * (TILE / 8))
        cov_blocks[((long)b * (is / 8) + (ty0 >> 3) + (t >> 1)) * W + (tx0 >> 5) + (t & 1)] = 0xf;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (aa) {
        const int R = is / 2, rtop = R - 1 - (ty0 >> 1);   // 32 x 32 outputs: one float4 per thread
        *reinterpret_cast<float4 *>(alpha + ((long)b * R + (rtop - (t >> 3))) * R + (tx0 >> 1) + 4 * (t & 7)) = z4;
    } else {
        float4 *p = reinterpret_cast<float4 *>(alpha + ((long)b * is + (is - 1 - ty0 - (t >> 4))) * is + tx0 + 4 * (t & 15));
        const long rs = 4 * (long)is;
        p[0] = z4; *(p - rs) = z4; *(p - 2 * rs) = z4; *(p - 3 * rs) = z4;
    }
}

#ifndef HM_FWD_MINB
#define HM_FWD_MINB 5
#endif
__global__ void __launch_bounds__(NTHREADS, HM_FWD_MINB)
raster_fwd_kernel(const FaceRec *__restrict__ recs, const FaceBox *__restrict__ boxes, int F, int is, int aa,
                  float near_, float far_, int32_t *__restrict__ face_index, float *__restrict__ alpha,
                  uint32_t *__restrict__ cov_row, uint32_t *__restrict__ cov_col, uint32_t *__restrict__ face_vis,
                  unsigned char *__restrict__ cov_blocks, const int *__restrict__ img_box) {
    extern __shared__ __align__(16) unsigned long long keys[];  // [TILE * TILE] z-buffer (dynamic: static + this > 48 KB)
    __shared__ int list[LISTCAP];
    __shared__ int cnt[2], next;
    __shared__ uint32_t roww[TILE][2];
    __shared__ uint32_t amb[TILE][2];   // ambiguous pixels (see above)
    __shared__ int n_amb;
    __shared__ unsigned hiz[HZN * HZN];  // per 4x4 block: largest winning depth so far (far when a pixel is empty)
    __shared__ __align__(4) unsigned short ambl[AMBCAP];
    __shared__ __align__(16) float4 wrec_all[NWARPS][8];  // per warp: the record of its current face
    __shared__ short spanx[NWARPS][64], spanpre[NWARPS][64];
    const int b = blockIdx.y;
    const int tiles_x = is / TILE;
    const int tx0 = (blockIdx.x % tiles_x) * TILE, ty0 = (blockIdx.x / tiles_x) * TILE;
    // ---- tiles no face touches (outside the image's pixel box, or an empty first scan of all faces) write constants
    bool untouched;
    {
        const int4 ib = __ldg(reinterpret_cast<const int4 *>(img_box) + b);   // {x0, y0, -x1, -y1}
        untouched = F == 0 || ib.x > tx0 + TILE - 1 || -ib.z < tx0 || ib.y > ty0 + TILE - 1 || -ib.w < ty0;
    }
    if (untouched) {
        write_untouched_tile(b, tx0, ty0, is, aa, face_index, alpha, cov_row, cov_col, cov_blocks);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    RowCtx rc;
    rc.is = is; rc.pow2 = (is & (is - 1)) == 0; rc.inv_is = 1.f / (float)is;
    const bool pow2 = rc.pow2;
    const float inv_is = rc.inv_is;
    const unsigned near_bits = __float_as_uint(near_), far_bits = __float_as_uint(far_);
    const unsigned long long empty = ((unsigned long long)far_bits << 32) | 0xffffffffull;
    recs += (long)b * F;
    boxes += (long)b * F;
    int base = 0;
    int n_first = next_batch(boxes, base, F, tx0, ty0, list, cnt, &next);
    if (n_first == 0 && base >= F) {
        write_untouched_tile(b, tx0, ty0, is, aa, face_index, alpha, cov_row, cov_col, cov_blocks);
        return;
    }
    for (int i = threadIdx.x; i < TILE * TILE / 2; i += NTHREADS)
        reinterpret_cast<ulonglong2 *>(keys)[i] = make_ulonglong2(empty, empty);
    if (threadIdx.x < 2 * TILE) (&amb[0][0])[threadIdx.x] = 0u;
    __syncthreads();

    float4 *wrec = wrec_all[warp];
    short *rowx = spanx[warp], *rowpre = spanpre[warp];
    bool first = true;
    while (first || base < F) {
        if (!first) next_batch(boxes, base, F, tx0, ty0, list, cnt, &next);
        first = false;
        // two passes: faces kept in their original winding (the outer layer of an outward-wound closed mesh)
        // first, the reversed copies second. Between them the winners are summarised per 4x4 block, so that a
        // hidden-layer face is dropped with one comparison per block instead of one per sample.
        bool filtered = false;
        for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            __syncthreads();
            hiz[threadIdx.x] = 0u;
            if (threadIdx.x == 0) next = 0;
            __syncthreads();
            for (int k = 0; k < TILE * TILE / NTHREADS; ++k) {  // 256 keys = 4 rows = one row of blocks, conflict-free
                unsigned m = (unsigned)(keys[k * NTHREADS + threadIdx.x] >> 32);
                m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
                if ((lane & 3) == 0) atomicMax(&hiz[k * HZN + (threadIdx.x % TILE) / HZ], m);
            }
            if (threadIdx.x == 0) n_amb = 0;   // (free until the passes end: counts the entries that survive)
            __syncthreads();
            // hidden-layer faces, thread per list entry: the interpolated depth is a convex combination of the corner
            // depths, so a face whose nearest corner (minus rounding slack) is behind the winner of every pixel of the
            // 4x4 blocks its box touches cannot win a pixel - dropped from the list here (one 16-byte load and a few
            // shared-memory reads per face) instead of by a whole warp in the face loop
            const int n1 = cnt[1];
            filtered = n1 >= HM_FWD_FILTER_MIN;   // (short lists: the barriers cost more than the warps' own test)
            for (int r0 = 0; filtered && r0 < n1; r0 += NTHREADS) {
                const int jj = r0 + threadIdx.x;
                int fce = -1;
                bool keep = false;
                if (jj < n1) {
                    fce = list[LISTCAP - 1 - jj];
                    const int4 q7 = __ldg(reinterpret_cast<const int4 *>(recs + fce) + 7);
                    const int X0 = max((int)(short)(q7.y & 0xffff), tx0), X1 = min((int)(short)(q7.z & 0xffff), tx0 + TILE - 1);
                    const int Y0 = max((int)(short)(q7.y >> 16), ty0), Y1 = min((int)(short)(q7.z >> 16), ty0 + TILE - 1);
                    if (X0 <= X1 && Y0 <= Y1) {
                        const float zmin = __int_as_float(q7.w);
                        const unsigned zmin_bits = zmin > 0.f ? __float_as_uint(zmin * (1.f - 1e-5f)) : 0u;
                        const int hx0 = (X0 - tx0) / HZ, hx1 = (X1 - tx0) / HZ, hy1 = (Y1 - ty0) / HZ;
                        for (int hy = (Y0 - ty0) / HZ; hy <= hy1 && !keep; ++hy)
                            for (int hx = hx0; hx <= hx1; ++hx)
                                if (!(zmin_bits > hiz[hy * HZN + hx])) { keep = true; break; }
                    }
                }
                __syncthreads();   // every entry of this round is read before the survivors are written over them
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                int at = 0;
                if (lane == 0 && km) at = atomicAdd(&n_amb, __popc(km));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (keep) list[LISTCAP - 1 - (at + __popc(km & ((1u << lane) - 1u)))] = fce;
                __syncthreads();
            }
            if (filtered && threadIdx.x == 0) cnt[1] = n_amb;
            __syncthreads();
        }
        const int np = cnt[pass];
        // ---- no barrier inside a pass: a warp draws one face of the pass at a time from a shared counter, and the
        //      128-byte record of its next face is fetched (one float4 in each of eight lanes) while the current one
        //      is processed. (With the records staged 64 at a time behind CTA barriers the warp with the largest
        //      faces of a chunk kept the other seven waiting: barrier stalls led the profile.)
        int j = 0;
        if (lane == 0) j = atomicAdd(&next, 1);
        j = __shfl_sync(FULL, j, 0);
        float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < np && lane < 8)
            pre = __ldg(reinterpret_cast<const float4 *>(recs + (pass ? list[LISTCAP - 1 - j] : list[j])) + lane);
        while (j < np) {
            __syncwarp();   // the previous face is done with wrec / rowx / rowpre
            if (lane < 8) wrec[lane] = pre;
            __syncwarp();
            {
                int jn = 0;
                if (lane == 0) jn = atomicAdd(&next, 1);
                j = __shfl_sync(FULL, jn, 0);
                if (j < np && lane < 8)
                    pre = __ldg(reinterpret_cast<const float4 *>(recs + (pass ? list[LISTCAP - 1 - j] : list[j])) + lane);
            }
            const float4 *rp = wrec;
            const float4 q2 = rp[2];
            const int4 q3 = *reinterpret_cast<const int4 *>(rp + 3);
            const float4 q0 = rp[0], q1 = rp[1];
            const float f0 = q0.x, f1 = q0.y, f2 = q0.z, f3 = q0.w, f4 = q1.x, f5 = q1.y, f6 = q1.z, f7 = q1.w,
                        f8 = q2.x;
            const int fn = __float_as_int(q2.y);
            const int4 q7 = *reinterpret_cast<const int4 *>(rp + 7);
            const int bx0 = (short)(q7.y & 0xffff), by0 = (short)(q7.y >> 16);   // forward box
            const int bx1 = (short)(q7.z & 0xffff), by1 = (short)(q7.z >> 16);
            const int X0 = max(bx0, tx0), X1 = min(bx1, tx0 + TILE - 1);
            const int Y0 = max(by0, ty0), Y1 = min(by1, ty0 + TILE - 1);
            const int w = X1 - X0 + 1, h = Y1 - Y0 + 1;
            if (w <= 0 || h <= 0) continue;
            if (q3.w & 1) {
                // degenerate face whose two windings are both front-facing: the reference also rasterises the
                // reversed copy F + f. Every pixel of its bounding box goes to the exact resolution below, which
                // handles both copies (such faces are rare and small).
                for (int r = lane; r < h; r += 32) {
                    const int yl = Y0 - ty0 + r, a = X0 - tx0, c = X1 - tx0;
                    const unsigned long long m = ((c == 63 ? ~0ull : ((1ull << (c + 1)) - 1ull)) >> a) << a;
                    if ((unsigned)m) atomicOr(&amb[yl][0], (unsigned)m);
                    if ((unsigned)(m >> 32)) atomicOr(&amb[yl][1], (unsigned)(m >> 32));
                }
            }
            const float e0x = f3 - f0, e0y = f4 - f1, e1x = f6 - f3, e1y = f7 - f4, e2x = f0 - f6, e2y = f1 - f7;
            // early z: the interpolated depth is a convex combination of the corner depths, so a face whose
            // nearest corner (minus rounding slack) is behind the current winner of a pixel cannot win it
            const float zmin = fminf(fminf(f2, f5), f8);
            const unsigned zmin_bits = zmin > 0.f ? __float_as_uint(zmin * (1.f - 1e-5f)) : 0u;
            if (pass == 1 && !filtered) {
                const int hx0 = (X0 - tx0) / HZ, hx1 = (X1 - tx0) / HZ, hy0 = (Y0 - ty0) / HZ, hy1 = (Y1 - ty0) / HZ;
                const int hx = hx0 + (lane & (HZN - 1));  // lanes: 16 block columns x 2 block rows
                bool vis = false;
                if (hx <= hx1)
                    for (int hy = hy0 + (lane >> 4); hy <= hy1; hy += 2) vis |= !(zmin_bits > hiz[hy * HZN + hx]);
                if (!__any_sync(FULL, vis)) continue;
            }
            // Covered pixels: per row of the box the exact interval of samples inside the triangle (clip_edge), then
            // the pixels of all rows flattened over the lanes of the group through a prefix sum of the interval
            // lengths - slivers and diagonal faces cost their area, not their bounding box, and no sample is tested twice.
            int n_px;
            {
                int len0 = 0, len1 = 0, xlo0 = X0, xlo1 = X0;
                const bool two = h > 32;   // (warp-uniform)
                for (int half = 0; half < (two ? 2 : 1); ++half) {
                    const int r = lane + 32 * half;
                    if (r < h) {
                        const int yi = Y0 + r;
                        const float yp = pow2 ? (float)(2 * yi + 1 - is) * inv_is : (float)(2 * yi + 1 - is) / (float)is;
                        int lo = X0, hi = X1;
                        clip_edge(rc, (yp - f1) * e0x, f0, e0y, X0, X1, lo, hi);
                        clip_edge(rc, (yp - f4) * e1x, f3, e1y, X0, X1, lo, hi);
                        clip_edge(rc, (yp - f7) * e2x, f6, e2y, X0, X1, lo, hi);
                        const int ln = max(hi - lo + 1, 0);
                        if (half == 0) { xlo0 = lo; len0 = ln; } else { xlo1 = lo; len1 = ln; }
                    }
                }
                int inc0 = len0, inc1 = len1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v0 = __shfl_up_sync(FULL, inc0, o), v1 = __shfl_up_sync(FULL, inc1, o);
                    if (lane >= o) { inc0 += v0; inc1 += v1; }
                }
                const int tot0 = __shfl_sync(FULL, inc0, 31), tot1 = __shfl_sync(FULL, inc1, 31);
                rowx[lane] = (short)xlo0; rowpre[lane] = (short)(inc0 - len0);
                rowx[lane + 32] = (short)xlo1; rowpre[lane + 32] = (short)(tot0 + inc1 - len1);
                n_px = tot0 + tot1;
                __syncwarp();
            }
            if (n_px == 0) continue;  // the box touches the tile, the triangle covers no sample of it
            // barycentric matrix in pixel coordinates and corner 1/z, prepared once per face by the setup kernel
            float inv[9], iz0, iz1, iz2;
            {
                const float4 i0 = rp[4], i1 = rp[5], i2 = rp[6];
                inv[0] = i0.x; inv[1] = i0.y; inv[2] = i0.z; inv[3] = i0.w;
                inv[4] = i1.x; inv[5] = i1.y; inv[6] = i1.z; inv[7] = i1.w;
                inv[8] = i2.x; iz0 = i2.y; iz1 = i2.z; iz2 = i2.w;
            }
            const int exact_face = q7.x;
            // a sample's depth lies between the corner depths (plus rounding): a face whose corners are clear of the
            // near / far planes needs no closeness test against them per sample
            const bool planes = exact_face || !(zmin > near_ * 1.0001f && fmaxf(fmaxf(f2, f5), f8) < far_ * 0.9999f);
            const short *gx = rowx, *gpre = rowpre;
            int row = 0, rend = 0, xoff = 0;   // current row, first pixel index of the next non-visited row, x - i of the row
            for (int i = lane; i < n_px; i += 32) {
                if (i == lane) {   // first pixel of the lane: binary search over the rows; afterwards the row only advances
                    if (h > 32 && gpre[32] <= i) row = 32;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1)
                        if (row + sft < h && gpre[row + sft] <= i) row += sft;
                    rend = row + 1 < h ? gpre[row + 1] : 0x7fffffff;
                    xoff = gx[row] - gpre[row];
                } else {
                    while (i >= rend) {
                        ++row;
                        xoff = gx[row] - rend;
                        rend = row + 1 < h ? gpre[row + 1] : 0x7fffffff;
                    }
                }
                const int xi = xoff + i, yi = Y0 + row;
                const int xl = xi - tx0, yl = yi - ty0;
                unsigned long long *kp = &keys[yl * TILE + xl];
                const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
                if (zmin_bits > (unsigned)(cur >> 32)) continue;
                float zp;
                if (exact_face) {
                    zp = exact_depth(inv, xi, yi, f2, f5, f8);
                } else {
                    float w0 = inv[0] * xi + inv[1] * yi + inv[2];
                    float w1 = inv[3] * xi + inv[4] * yi + inv[5];
                    float w2 = inv[6] * xi + inv[7] * yi + inv[8];
                    w0 = fminf(fmaxf(w0, 0.f), 1.f);
                    w1 = fminf(fmaxf(w1, 0.f), 1.f);
                    w2 = fminf(fmaxf(w2, 0.f), 1.f);
                    zp = __fdividef(w0 + w1 + w2, w0 * iz0 + w1 * iz1 + w2 * iz2);
                }
                const unsigned zb = __float_as_uint(zp);
                bool flag = planes && (ulp_close(zb, near_bits) || ulp_close(zb, far_bits));
                if (zp > near_ && zp < far_) {  // also rejects NaN
                    const unsigned long long key = ((unsigned long long)zb << 32) | (unsigned)fn;
                    unsigned long long seen = cur;
                    if (key < cur) seen = atomicMin(kp, key);
                    flag |= ((unsigned)seen != 0xffffffffu) && ulp_close(zb, (unsigned)(seen >> 32));
                }
                if (flag) atomicOr(&amb[yl][xl >> 5], 1u << (xl & 31));
            }
        }
        }
    }

    // ---- ambiguous pixels: start over with the reference arithmetic, against every face of the image
    __syncthreads();
    const int any_amb = __syncthreads_or(threadIdx.x < 2 * TILE && (&amb[0][0])[threadIdx.x] != 0u);
    for (; any_amb;) {
        __syncthreads();
        if (threadIdx.x == 0) n_amb = 0;
        __syncthreads();
        if (threadIdx.x < 2 * TILE) {
            unsigned word = (&amb[0][0])[threadIdx.x];
            while (word) {
                const int bit = __ffs(word) - 1;
                const int pos = atomicAdd(&n_amb, 1);
                if (pos >= AMBCAP) break;
                word &= word - 1;
                const int yl = threadIdx.x >> 1, xl = (threadIdx.x & 1) * 32 + bit;
                ambl[pos] = (unsigned short)(xl | (yl << 6));
                keys[yl * TILE + xl] = empty;
            }
            (&amb[0][0])[threadIdx.x] = word;
        }
        __syncthreads();
        const int total = n_amb;
        const int na = min(total, AMBCAP);
        if (na == 0) break;
        int base2 = 0;
        while (base2 < F) {
            const int n = next_batch(boxes, base2, F, tx0, ty0, list, cnt, &next);
            const int n_front = cnt[0];
            for (int li = threadIdx.x >> 5; li < n; li += NWARPS) {
                const FaceRec *rp = recs + batch_face(list, n_front, li);
                int fn = __ldg(reinterpret_cast<const int *>(rp) + 9);
                if (fn < 0) continue;
                float f[9], inv[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) { f[k] = __ldg(reinterpret_cast<const float *>(rp) + k); inv[k] = __ldg(reinterpret_cast<const float *>(rp) + 16 + k); }
                const int q = __ldg(reinterpret_cast<const int *>(rp) + 13), q2 = __ldg(reinterpret_cast<const int *>(rp) + 14);
                const int bx0 = (short)(q & 0xffff), by0 = (short)(q >> 16), bx1 = (short)(q2 & 0xffff), by1 = (short)(q2 >> 16);
                const int copies = (__ldg(reinterpret_cast<const int *>(rp) + 15) & 1) ? 2 : 1;
#pragma unroll 1
                for (int copy = 0; copy < copies; ++copy) {
                    if (copy == 1) {  // the reversed copy F + f of a degenerate face
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            float t = f[k]; f[k] = f[6 + k]; f[6 + k] = t;
                            t = inv[k]; inv[k] = inv[6 + k]; inv[6 + k] = t;
                        }
                        fn += F;
                    }
                    for (int j = lane; j < na; j += 32) {
                        const unsigned e = ambl[j];
                        const int xl = e & 63u, yl = e >> 6;
                        const int xi = tx0 + xl, yi = ty0 + yl;
                        if (xi < bx0 || xi > bx1 || yi < by0 || yi > by1) continue;
                        const float yp = pow2 ? (float)(2 * yi + 1 - is) * inv_is : (float)(2 * yi + 1 - is) / (float)is;
                        const float xp = pow2 ? (float)(2 * xi + 1 - is) * inv_is : (float)(2 * xi + 1 - is) / (float)is;
                        if ((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) continue;
                        if ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) continue;
                        if ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])) continue;
                        const float zp = exact_depth(inv, xi, yi, f[2], f[5], f[8]);
                        if (zp > near_ && zp < far_)
                            atomicMin(&keys[yl * TILE + xl], ((unsigned long long)__float_as_uint(zp) << 32) | (unsigned)fn);
                    }
                }
            }
        }
        if (total <= AMBCAP) break;
    }
    __syncthreads();

    // ---- write-out: face_index rows (four pixels per thread, 16-byte stores), coverage words, and the faces that own
    //      a pixel (face_vis: one bit per face of the doubled numbering; collected in shared memory - the face list
    //      is free now - when it fits, one global atomicOr per non-zero word and tile)
    const int nvw = (2 * F + 31) / 32;
    const bool vis_shared = nvw <= LISTCAP;
    uint32_t *vis_out = face_vis ? face_vis + (long)b * nvw : nullptr;
    if (vis_out && vis_shared) {
        for (int i = threadIdx.x; i < nvw; i += NTHREADS) list[i] = 0;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < TILE * TILE / 4; i += NTHREADS) {
        const int yl = i / (TILE / 4), x4 = i % (TILE / 4);  // a warp covers two rows of the tile
        const ulonglong2 k01 = *reinterpret_cast<const ulonglong2 *>(&keys[4 * i]);
        const ulonglong2 k23 = *reinterpret_cast<const ulonglong2 *>(&keys[4 * i + 2]);
        const int4 fi4 = make_int4((int)(unsigned)k01.x, (int)(unsigned)k01.y, (int)(unsigned)k23.x, (int)(unsigned)k23.y);
        *reinterpret_cast<int4 *>(face_index + ((long)b * is + (ty0 + yl)) * is + tx0 + 4 * x4) = fi4;  // empty = -1
        if (vis_out) {   // one mark per run of equal owners along the row
            int prev = __shfl_up_sync(0xffffffffu, fi4.w, 1);
            if ((lane & 15) == 0) prev = -1;
            const int o[4] = {fi4.x, fi4.y, fi4.z, fi4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (o[k] >= 0 && o[k] != prev) {
                    if (vis_shared) atomicOr(reinterpret_cast<unsigned *>(&list[o[k] >> 5]), 1u << (o[k] & 31));
                    else atomicOr(&vis_out[o[k] >> 5], 1u << (o[k] & 31));
                }
                prev = o[k];
            }
        }
        unsigned nib = (fi4.x != -1 ? 1u : 0u) | (fi4.y != -1 ? 2u : 0u) | (fi4.z != -1 ? 4u : 0u) | (fi4.w != -1 ? 8u : 0u);
        nib <<= 4 * (lane & 7);  // eight consecutive lanes make one 32-pixel word
        nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
        if ((lane & 7) == 0) {
            const int w = x4 >> 3;
            roww[yl][w] = nib;
            if (cov_row) cov_row[((long)b * is + (ty0 + yl)) * (is / 32) + (tx0 >> 5) + w] = nib;
        }
    }
    __syncthreads();
    if (vis_out && vis_shared)
        for (int i = threadIdx.x; i < nvw; i += NTHREADS) {
            const unsigned w = (unsigned)list[i];
            if (w) atomicOr(&vis_out[i], w);
        }
    if (cov_blocks && threadIdx.x < 2 * (TILE / 8)) {
        // 8x8-block summary of the coverage: bit j of cov_blocks[band][w] = block 4w + j of the 8-row band has an
        // uncovered pixel (the backward's test for faces that can have in-sweeps)
        const int band = threadIdx.x >> 1, w = threadIdx.x & 1;
        unsigned a = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < 8; ++r) a &= roww[band * 8 + r][w];
        a = ~a;
        cov_blocks[((long)b * (is / 8) + (ty0 >> 3) + band) * (is / 32) + (tx0 >> 5) + w] =
            (unsigned char)(((a & 0xffu) ? 1u : 0u) | ((a & 0xff00u) ? 2u : 0u) | ((a & 0xff0000u) ? 4u : 0u) |
                            ((a & 0xff000000u) ? 8u : 0u));
    }
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
// Streaming kernel, one thread per group of output words, no shared memory: the first n_row threads of an image
// produce row words (coalesced reads of grad_alpha rows), the others column words (coalesced across columns).
// With anti-aliasing an output pixel (r, c) is the 2x2 raster block rows is-1-2r, is-2-2r / columns 2c, 2c+1,
// so 16 gradient signs expand to one 32-bit word of two raster lines.
__device__ __forceinline__ unsigned dup_bits16(unsigned v) {  // bit k -> bits 2k, 2k + 1
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v | (v << 1);
}

template <bool AA>
__global__ void __launch_bounds__(NTHREADS)
grad_prep_kernel(const float *__restrict__ grad_alpha, const uint32_t *__restrict__ cov_row,
                 const uint32_t *__restrict__ cov_col, int is, uint32_t *__restrict__ m_row,
                 uint32_t *__restrict__ m_col) {
    const int b = blockIdx.y;
    const int W = is / 32, R = AA ? is / 2 : is;
    const int n_half = R * W;  // row groups, then as many column groups
    const int idx = blockIdx.x * NTHREADS + threadIdx.x;
    if (idx >= 2 * n_half) return;
    const float *g = grad_alpha + (long)b * R * R;
    const long plane = (long)is * W;
    if (idx < n_half) {
        // ---- row words: output row r, word w
        const int r = idx / W, w = idx % W;
        unsigned neg = 0, pos = 0;
        if (AA) {
            const float4 *p = reinterpret_cast<const float4 *>(g + (long)r * R + 16 * w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v = __ldg(p + k);
                const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    neg |= (a[j] < 0.f ? 1u : 0u) << (4 * k + j);
                    pos |= (a[j] > 0.f ? 1u : 0u) << (4 * k + j);
                }
            }
            neg = dup_bits16(neg); pos = dup_bits16(pos);
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int y = is - 1 - 2 * r - d;
                const long o = ((long)b * 2 * is + y) * W + w;
                const unsigned A = __ldg(cov_row + ((long)b * is + y) * W + w);
                m_row[o] = neg & ~A;
                m_row[o + plane] = pos & A;
            }
        } else {
            const float4 *p = reinterpret_cast<const float4 *>(g + (long)r * R + 32 * w);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 v = __ldg(p + k);
                const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    neg |= (a[j] < 0.f ? 1u : 0u) << (4 * k + j);
                    pos |= (a[j] > 0.f ? 1u : 0u) << (4 * k + j);
                }
            }
            const int y = is - 1 - r;
            const long o = ((long)b * 2 * is + y) * W + w;
            const unsigned A = __ldg(cov_row + ((long)b * is + y) * W + w);
            m_row[o] = neg & ~A;
            m_row[o + plane] = pos & A;
        }
    } else {
        // ---- column words: output column c, word w (bit j <-> raster row y = 32 w + j); adjacent threads take
        //      adjacent columns, so the strided reads down a column coalesce across the warp
        const int k2 = idx - n_half;
        const int w = k2 / R, c = k2 % R;
        unsigned neg = 0, pos = 0;
        if (AA) {
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const float v = __ldg(g + (long)(R - 1 - 16 * w - k) * R + c);  // rows y = 32 w + 2k, 32 w + 2k + 1
                neg |= (v < 0.f ? 1u : 0u) << k;
                pos |= (v > 0.f ? 1u : 0u) << k;
            }
            neg = dup_bits16(neg); pos = dup_bits16(pos);
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int x = 2 * c + d;
                const long o = ((long)b * 2 * is + x) * W + w;
                const unsigned A = __ldg(cov_col + ((long)b * is + x) * W + w);
                m_col[o] = neg & ~A;
                m_col[o + plane] = pos & A;
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                const float v = __ldg(g + (long)(is - 1 - 32 * w - j) * R + c);
                neg |= (v < 0.f ? 1u : 0u) << j;
                pos |= (v > 0.f ? 1u : 0u) << j;
            }
            const long o = ((long)b * 2 * is + c) * W + w;
            const unsigned A = __ldg(cov_col + ((long)b * is + c) * W + w);
            m_col[o] = neg & ~A;
            m_col[o + plane] = pos & A;
        }
    }
}

// ------------------------------------------------------------------------------------------ sweep runs
// Run-length form of the four sweep masks: maximal stretches of consecutive set bits of one line that
// carry the same |grad| (the silhouette loss gradient is piecewise constant: 2 * norm * (rend - ref) takes a
// handful of values and is constant over whole missing / excess regions). A sweep then costs O(runs on the
// line) instead of O(set pixels). Lines with more than HM_RASTER_RUN_CAP runs keep the bit-line path.
//   runs       [B][4][is][HM_RASTER_RUN_CAP] uint2 {start | end << 16, |grad| bits}
//   run_counts [B][4][is]  count (15 = overflow) | first set pixel << 4 | last set pixel << 16
//   list 0 mn_row, 1 mp_row, 2 mn_col, 3 mp_col
constexpr int RCAP = HM_RASTER_RUN_CAP;
constexpr unsigned RUN_OVERFLOW = 15u;
static_assert(RCAP < 15, "run count field");

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
    int n = 0, rs = -1, re = -1, first = 0;
    float rg = 0.f;
    for (int w = 0; w < W; ++w) {
        unsigned bits = words[w];
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            const int d1 = w * 32 + bit;
            if (rs < 0) first = d1;
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
    run_counts[((long)b * 4 + l4) * is + line] =
        (n > RCAP ? RUN_OVERFLOW : (unsigned)n) | ((unsigned)first << 4) | ((unsigned)max(re, 0) << 16);
}

// ------------------------------------------------------------------------------------------ backward
// @region eval_item
// One MUFU.RCP (1 ulp) instead of the IEEE reciprocal sequence: the sweep sums tolerate it (parity bar 1e-4).
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// psi(z2) - psi(z1) = sum_{k=0}^{n-1} 1 / (z1 + k) for z2 = z1 + n, z1 >= 4 (asymptotic series, error < 1e-7).
// The logarithm is MUFU.LG2 of 1 + n / z1: absolute error ~1e-7, against a sum whose leading terms are O(1 / z).
__device__ __forceinline__ float harmonic_span(float z1, float n) {
    const float z2 = z1 + n;
    const float i1 = rcp_fast(z1), i2 = rcp_fast(z2);
    const float a1 = i1 * i1, a2 = i2 * i2;
    float r = __logf(1.f + n * i1);
    r += 0.5f * (i1 - i2);
    r += (1.f / 12.f) * (a1 - a2);
    r -= (1.f / 120.f) * (a1 * a1 - a2 * a2);
    r += (1.f / 252.f) * (a1 * a1 * a1 - a2 * a2 * a2);
    return r;
}

// One (crossing, run) work item: the part [s, e] of one run of one sweep line seen from the crossing at
// x = d1_cross of scan-line d0. A sweep runs away from the crossing, so the whole item lies on one side of it:
// pixel k of the item (counted from the crossing) sits at distance z + k. The first NEAR_N pixels (the large terms)
// are evaluated with the reference's expression; beyond, every pixel has the same weight G and the sum of
// 1 / dist is the harmonic sum G / K * sum 1 / (z + k + eps / |K|), taken in closed form (dist = K (d1 - x) +- eps,
// K = c * 2 / is). Straight-line code: every lane of a warp does the same work whatever its item looks like.
#ifndef HM_NEAR_N
#define HM_NEAR_N 4
#endif
constexpr int NEAR_N = HM_NEAR_N;
__device__ __forceinline__ void eval_item(float x, float c0, float c1, float G, int s, int e, bool has0, bool has1,
                                          float inv_is2, float eps, float &a0, float &a1) {
    const float K0 = c0 * inv_is2, K1 = c1 * inv_is2;
    const float rK0 = rcp_fast(K0), rK1 = rcp_fast(K1);
    const float del0 = eps * fabsf(rK0), del1 = eps * fabsf(rK1);
    const bool left = (float)e <= x;
    const float sgn = left ? -1.f : 1.f;
    const float z = left ? x - (float)e : (float)s - x;  // distance of the item's nearest pixel (>= 0)
    const int n = e - s + 1;
    float h0 = 0.f, h1 = 0.f;
#pragma unroll
    for (int k = 0; k < NEAR_N; ++k) {
        const float dd = sgn * (z + (float)k);
        float dist0 = K0 * dd, dist1 = K1 * dd;
        dist0 = (0.f < dist0) ? dist0 + eps : dist0 - eps;
        dist1 = (0.f < dist1) ? dist1 + eps : dist1 - eps;
        const float t0 = rcp_fast(dist0), t1 = rcp_fast(dist1);
        if (k < n) { h0 += t0; h1 += t1; }
    }
    const float nf = (float)max(n - NEAR_N, 0), zf = z + (float)NEAR_N;
    h0 += sgn * harmonic_span(zf + del0, nf) * rK0;
    h1 += sgn * harmonic_span(zf + del1, nf) * rK1;
    a0 = has0 ? -G * h0 : 0.f;
    a1 = has1 ? -G * h1 : 0.f;
}

// @region sweep_bits
// Bit-line walk for lines whose run list overflowed: visits the set bits of `line` in [a, c].
__device__ __noinline__ void sweep_bits(const uint32_t *line, int a, int c, int axis, int d0, float d1_cross,
                                           float c0, float c1, bool has0, bool has1, const BwdCtx &ctx, float &acc0,
                                           float &acc1) {
    if (a > c) return;
    const int wa = a >> 5, wc = c >> 5;
    for (int w = wa; w <= wc; ++w) {
        unsigned bits = __ldg(line + w);
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

// ------------------------------------------------------------------------------------------ backward geometry
// @region geom
struct TaskGeom {
    float p0d0, p0d1, p1d0, p2d0, p2d1, slope, s02, s21, ka;
    int dir, d0_from, d0_to, vid0, vid1;
};

__device__ __forceinline__ float sel3(const float (&a)[3], int i) { return i == 0 ? a[0] : i == 1 ? a[1] : a[2]; }
// The quotient of an edge walked backwards: (-a) / (-b) rounds like a / b, except that b = +0 both ways (x - x = +0),
// so an infinite slope changes sign.
__device__ __forceinline__ float backwards(float slope) { return fabsf(slope) == INFINITY ? -slope : slope; }

// The backward record of a face, optionally as its reversed (fill_back) copy.
struct BwdFace {
    float px[3], py[3], sl[3][2];
    int v[3], fn;
    bool both, irregular;
};
__device__ __forceinline__ BwdFace load_bwd_face(const BwdRec *rp) {
    BwdFace f;
    const float4 q0 = __ldg(reinterpret_cast<const float4 *>(rp)), q1 = __ldg(reinterpret_cast<const float4 *>(rp) + 1),
                 q2 = __ldg(reinterpret_cast<const float4 *>(rp) + 2);
    const int4 q3 = __ldg(reinterpret_cast<const int4 *>(rp) + 3);
    f.px[0] = q0.x; f.px[1] = q0.y; f.px[2] = q0.z; f.py[0] = q0.w; f.py[1] = q1.x; f.py[2] = q1.y;
    f.sl[0][0] = q1.z; f.sl[0][1] = q1.w; f.sl[1][0] = q2.x; f.sl[1][1] = q2.y; f.sl[2][0] = q2.z; f.sl[2][1] = q2.w;
    f.v[0] = q3.x; f.v[1] = q3.y; f.v[2] = q3.z;
    f.fn = q3.w < 0 ? -1 : (q3.w & BWD_FN_MASK);
    f.both = q3.w >= 0 && (q3.w & BWD_BOTH);
    f.irregular = q3.w >= 0 && (q3.w & BWD_IRREGULAR);
    return f;
}
__device__ __forceinline__ void reverse_bwd_face(BwdFace &f, int F) {  // corners 0 and 2 trade places, fn -> F + f
    float t = f.px[0]; f.px[0] = f.px[2]; f.px[2] = t;
    t = f.py[0]; f.py[0] = f.py[2]; f.py[2] = t;
    t = f.sl[0][0]; f.sl[0][0] = f.sl[1][0]; f.sl[1][0] = t;   // reversed edge 0 = old edge 1 walked backwards
    t = f.sl[0][1]; f.sl[0][1] = f.sl[1][1]; f.sl[1][1] = t;
#pragma unroll
    for (int e = 0; e < 3; ++e) { f.sl[e][0] = backwards(f.sl[e][0]); f.sl[e][1] = backwards(f.sl[e][1]); }
    const int v = f.v[0]; f.v[0] = f.v[2]; f.v[2] = v;
    f.fn += F;
}
__device__ __forceinline__ TaskGeom task_geom(const BwdFace &f, int e, int axis, int is) {
    TaskGeom g;
    const int e1 = e == 2 ? 0 : e + 1, e2 = e == 0 ? 2 : e - 1;
    // (d0, d1) = (x, y) for axis 0, (y, x) for axis 1   (selects, not indexing: the record stays in registers)
    const float x0 = sel3(f.px, e), y0 = sel3(f.py, e), x1 = sel3(f.px, e1), y1 = sel3(f.py, e1),
                x2 = sel3(f.px, e2), y2 = sel3(f.py, e2);
    g.p0d0 = axis ? y0 : x0; g.p0d1 = axis ? x0 : y0;
    g.p1d0 = axis ? y1 : x1;
    g.p2d0 = axis ? y2 : x2; g.p2d1 = axis ? x2 : y2;
    const float s0 = axis ? f.sl[0][1] : f.sl[0][0], s1 = axis ? f.sl[1][1] : f.sl[1][0], s2 = axis ? f.sl[2][1] : f.sl[2][0];
    g.slope = e == 0 ? s0 : e == 1 ? s1 : s2;
    // edge e2 runs corner e2 -> e: the reference's (p2 - p0) quotient walked backwards; edge e1 runs e1 -> e2: (p1 - p2)
    g.s02 = backwards(e == 0 ? s2 : e == 1 ? s0 : s1);
    g.s21 = backwards(e == 0 ? s1 : e == 1 ? s2 : s0);
    g.vid0 = e == 0 ? f.v[0] : e == 1 ? f.v[1] : f.v[2];
    g.vid1 = e == 0 ? f.v[1] : e == 1 ? f.v[2] : f.v[0];
    if (axis == 0) g.dir = (g.p0d0 < g.p1d0) ? -1 : 1;
    else g.dir = (g.p0d0 < g.p1d0) ? 1 : -1;
    g.d0_from = __float2int_rz(fmaxf(ceilf(fminf(g.p0d0, g.p1d0)), 0.f));
    g.d0_to = __float2int_rz(fminf(fmaxf(g.p0d0, g.p1d0), (float)(is - 1)));
    g.ka = g.p1d0 - g.p0d0;
    return g;
}

struct SweepSrc {
    const uint2 *runs;            // [4][is][RCAP] of this image
    const uint32_t *run_info;     // [4][is] of this image
    const uint32_t *m_row, *m_col;  // bit lines of this image [2][is][W]
    int is, W;
    float inv_is2, eps;
    BwdCtx ctx;
};
__device__ __forceinline__ SweepSrc sweep_src(int b, int is, int aa, float eps, const float *grad_alpha,
                                              const uint32_t *m_row, const uint32_t *m_col, const uint2 *runs,
                                              const uint32_t *run_info) {
    SweepSrc S;
    const int W = is / 32;
    S.runs = runs + (long)b * 4 * is * RCAP;
    S.run_info = run_info + (long)b * 4 * is;
    S.m_row = m_row + (long)b * 2 * is * W;
    S.m_col = m_col + (long)b * 2 * is * W;
    S.is = is; S.W = W; S.inv_is2 = 2.f / (float)is; S.eps = eps;
    S.ctx.is = is; S.ctx.aa = aa; S.ctx.R = aa ? is / 2 : is; S.ctx.eps = eps;
    S.ctx.grad = grad_alpha + (long)b * S.ctx.R * S.ctx.R;
    return S;
}

// @region bwd2
// ------------------------------------------------------------------------------------------ backward, segment kernel
// backward_pixel_map as the reference organises it - a walk over the scan-lines of every (face, edge, axis) task with
// the ownership test on face_index - cut into units that fill a GPU: one THREAD per segment of <= B2_SEG consecutive
// scan-lines of one task. A CTA takes rounds of NTHREADS faces of one image: (1) a thread per face counts the segments
// of its six tasks (both copies of a both-windings face) and decides whether the face can have in-sweeps at all (an
// uncovered pixel near its pixel bounding box, looked up in an 8x8-block summary of the coverage built once per CTA,
// or a corner on an integer pixel coordinate); a block scan turns the counts into a shared work list; (2) a thread per
// list entry: the face_index reads of its scan-lines are issued together, then each crossing that owns its in-pixel
// sweeps the missing-coverage runs of its line beyond the crossing, and crossings of silhouette faces sweep the span of
// the triangle; (crossing, run) items are evaluated in place (eval_item_fast), sums stay in two registers and leave as
// one atomicAdd per segment and edge vertex into grad_ndc (the vertices_to_faces scatter-add is fused).
#ifndef HM_BWD_SEG
#define HM_BWD_SEG 8
#endif
constexpr int B2_SEG = HM_BWD_SEG;       // scan-lines per segment (a multiple of 4, <= 8)
constexpr int B2_CAP = 2048;            // list entries per slice
#ifndef HM_BWD_SMALL_MESH
#define HM_BWD_SMALL_MESH 768            // meshes up to this many faces take half-size rounds
#endif

__device__ __noinline__ void sweep_walk(const SweepSrc &S, int ls, int d0, int ra, int rc, float x, float K0, float K1,
                                        unsigned flags, float &a0, float &a1) {
    const int col = ls >> 1;
    const uint32_t *line = (col ? S.m_col : S.m_row) + ((long)(ls & 1) * S.is + d0) * S.W;
    const float is2 = 0.5f * (float)S.is;   // sweep_bits takes the coefficients before the 2 / is scale
    sweep_bits(line, ra, rc, col ? 0 : 1, d0, x, K0 * is2, K1 * is2, flags & 1u, flags & 2u, S.ctx, a0, a1);
}

// Per-warp queue of the sweeps that found runs, in pair order: x, K0, K1 (crossing position, distance coefficients
// with 2 / is included), meta = has0 | has1 << 1 | segment (lane of the warp pass) << 2 | first run << 7 | runs << 10,
// where = ra | rc << 10 | (list * is + scan-line) << 20.
constexpr int B2_QCAP = 64;   // ring: a leftover of < 32 plus one push of <= 32
struct SweepQueue2 {
    float x[B2_QCAP], K0[B2_QCAP], K1[B2_QCAP];
    unsigned meta[B2_QCAP], where[B2_QCAP];
};
// Evaluates sweeps [head, head + n) of the ring (n <= 32, one per lane): the runs of a sweep one after the other
// (straight-line item code; the queue holds sweeps with runs only, most have one or two). The sweeps of one segment
// are neighbours: segmented sum over the lanes, the first lane of every group adds to the segment's accumulator
// (batches run one after the other: a plain read-modify-write, and the order of the sum is fixed).
__device__ __noinline__ void b2_drain(const SweepQueue2 &sq, int head, int n, const uint2 *__restrict__ runs, float eps,
                                      float (*sacc)[2]) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float a0 = 0.f, a1 = 0.f;
    int owner = 32 + lane, nr = 0;
    float x = 0.f, K0 = 0.f, K1 = 0.f;
    unsigned meta = 0, where = 0;
    if (lane < n) {
        const int s_ = (head + lane) & (B2_QCAP - 1);
        x = sq.x[s_]; K0 = sq.K0[s_]; K1 = sq.K1[s_]; meta = sq.meta[s_]; where = sq.where[s_];
        owner = (meta >> 2) & 31;
        nr = meta >> 10;
    }
    const uint2 *rl = runs + (long)(where >> 20) * RCAP + ((meta >> 7) & 7u);
    const int ra = where & 0x3ffu, rc = (where >> 10) & 0x3ffu;
    const int rounds = __reduce_max_sync(FULL, nr);
    for (int k = 0; k < rounds; ++k) {
        if (k < nr) {
            const uint2 run = __ldg(rl + k);
            float t0, t1;
            eval_item(x, K0, K1, __uint_as_float(run.y), max(ra, (int)(run.x & 0xffffu)), min(rc, (int)(run.x >> 16)),
                      meta & 1u, meta & 2u, 1.f, eps, t0, t1);
            a0 += t0; a1 += t1;
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v0 = __shfl_down_sync(FULL, a0, o), v1 = __shfl_down_sync(FULL, a1, o);
        const int ow = __shfl_down_sync(FULL, owner, o);
        if (lane + o < 32 && ow == owner) { a0 += v0; a1 += v1; }
    }
    const int prev_owner = __shfl_up_sync(FULL, owner, 1);
    if (owner < 32 && (lane == 0 || prev_owner != owner)) {
        sacc[owner][0] += a0;
        sacc[owner][1] += a1;
    }
    __syncwarp();
}

#ifndef HM_BWD_MINB
#define HM_BWD_MINB 4
#endif
// Geometry of the 32 tasks a warp is working on, by lane (structure of arrays: the consumer of a crossing reads the
// fields of the lane that produced it).
struct TaskStash {
    float slope[32], p0d0[32], p0d1[32], p1d0[32], p2d0[32], p2d1[32], s02[32], s21[32];
    int misc[32];   // d0a | axis << 12 | dir > 0 << 13
};
constexpr int B2_PAIRS = 32 * 2 * B2_SEG;   // (lane, scan-line, kind) of the crossings worth a sweep: at most two per line

__global__ void __launch_bounds__(NTHREADS, HM_BWD_MINB)
raster_bwd_kernel(const BwdRec *__restrict__ brecs, const FaceBox *__restrict__ boxes, int F, int V, int is, int aa,
                   float eps, const int32_t *__restrict__ face_index, const float *__restrict__ grad_alpha,
                   const uint32_t *__restrict__ cov_row, const uint32_t *__restrict__ cov_col,
                   const uint32_t *__restrict__ face_vis, const unsigned char *__restrict__ cov_blocks,
                   const uint32_t *__restrict__ m_row, const uint32_t *__restrict__ m_col,
                   const uint2 *__restrict__ runs, const uint32_t *__restrict__ run_info, float *__restrict__ grad_ndc,
                   unsigned long long *__restrict__ grad_fixed, int fpr) {
    __shared__ uint32_t list[B2_CAP];
    __shared__ int wsum[NWARPS];
    __shared__ SweepSrc S;
    __shared__ float sacc[NWARPS][32][2];
    __shared__ SweepQueue2 sq2[NWARPS];
    __shared__ TaskStash stash[NWARPS];
    __shared__ unsigned short pairs[NWARPS][B2_PAIRS];
    const int b = blockIdx.y;
    const int W = is / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    brecs += (long)b * F;
    boxes += (long)b * F;
    grad_ndc += (long)b * V * 3;
    if (grad_fixed) grad_fixed += (long)b * V * 3;
    face_index += (long)b * is * is;
    cov_row += (long)b * is * W;
    cov_col += (long)b * is * W;
    face_vis += (long)b * HM_FACE_VIS_WORDS(F);
    cov_blocks += (long)b * (is / 8) * W;
    if (threadIdx.x == 0) S = sweep_src(b, is, aa, eps, grad_alpha, m_row, m_col, runs, run_info);
    // fpr = faces per round (<= NTHREADS, thread per face in phase 1): small meshes take rounds of 128 faces so that
    // the launch has enough CTAs for even waves (500 faces x 480 images: 960 CTAs of 256 faces = 1.6 waves)
    const int n_rounds = (F + fpr - 1) / fpr;
    for (int round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        // ---- (1) thread per face: is it visible (owns a pixel: out-sweeps), can it have in-sweeps (an uncovered pixel
        //      in the 8x8 blocks its pixel bounding box touches, or irregular), how many segments do its tasks have
        const int f = (int)threadIdx.x < fpr ? round * fpr + (int)threadIdx.x : F;
        unsigned ns = 0, ns_hi = 0;   // segments per task, 8 bits each
        int n_f = 0;
        unsigned fflags = 0;          // 1: boundary, 2: first copy visible, 4: there is a second copy, 8: second copy visible
        if (f < F) {
            const int4 q3 = __ldg(reinterpret_cast<const int4 *>(brecs + f) + 3);
            if (q3.w >= 0) {
                const int fn = q3.w & BWD_FN_MASK;
                const uint4 s4 = __ldg(reinterpret_cast<const uint4 *>(brecs + f) + 4);
                const uint2 s5 = __ldg(reinterpret_cast<const uint2 *>(brecs + f) + 10);
                // irregular faces and faces with a steep task (sweep ends extrapolated with large slopes) always sweep
                bool boundary = (q3.w & BWD_IRREGULAR) || ((s4.x | s4.y | s4.z | s4.w | s5.x | s5.y) & (1u << 25));
                if (!boundary) {
                    FaceBox bx = boxes[f];   // clamped pixel bbox with one pixel of slack
                    bx.y1 &= FBOX_MASK;
                    const int w0 = bx.x0 >> 5, w1 = bx.x1 >> 5;
                    for (int band = bx.y0 >> 3; band <= (bx.y1 >> 3) && !boundary; ++band)
                        for (int w = w0; w <= w1; ++w) {
                            const int lo = max((bx.x0 >> 3) - 4 * w, 0), hi = min((bx.x1 >> 3) - 4 * w, 3);
                            if (__ldg(cov_blocks + band * W + w) & ((0xfu >> (3 - hi)) & (0xfu << lo))) { boundary = true; break; }
                        }
                    if (boundary) {   // some 8x8 block the box touches has a hole: look at the box itself
                        boundary = false;
                        for (int y = bx.y0; y <= bx.y1 && !boundary; ++y)
                            for (int w = w0; w <= w1; ++w) {
                                const int lo = max(bx.x0 - 32 * w, 0), hi = min(bx.x1 - 32 * w, 31);
                                if (~__ldg(cov_row + y * W + w) & ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo))) { boundary = true; break; }
                            }
                    }
                }
                const bool vis0 = (__ldg(face_vis + (fn >> 5)) >> (fn & 31)) & 1u;
                const bool both = q3.w & BWD_BOTH;
                const bool vis1 = both && ((__ldg(face_vis + ((fn + F) >> 5)) >> ((fn + F) & 31)) & 1u);
                fflags = (boundary ? 1u : 0u) | (vis0 ? 2u : 0u) | (both ? 4u : 0u) | (vis1 ? 8u : 0u);
                const int copies = ((boundary || vis0) ? 1 : 0) + ((both && (boundary || vis1)) ? 1 : 0);
                if (copies) {
                    const unsigned sp[6] = {s4.x, s4.y, s4.z, s4.w, s5.x, s5.y};
                    int tot = 0;
#pragma unroll
                    for (int t = 0; t < 6; ++t) {
                        const int len = (int)((sp[t] >> 12) & 0xfffu) - (int)(sp[t] & 0xfffu) + 1;
                        const int n = len > 0 ? (len + B2_SEG - 1) / B2_SEG : 0;   // <= 128
                        if (t < 4) ns |= (unsigned)n << (8 * t); else ns_hi |= (unsigned)n << (8 * (t - 4));
                        tot += n;
                    }
                    n_f = tot * copies;
                }
            }
        }
        // block-wide exclusive scan of n_f
        int incl = n_f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        __syncthreads();   // previous round's list fully consumed, wsum free
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = incl - n_f, total = 0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            const int v = wsum[w];
            if (w < warp) base += v;
            total += v;
        }
        for (int lo = 0; lo < total; lo += B2_CAP) {
            if (lo) __syncthreads();
            // ---- entries of this slice: face (local) | task << 8 | copy << 11 | boundary << 12 | visible << 13 | segment << 14
            if (n_f && base < lo + B2_CAP && base + n_f > lo) {
                int idx = base;
                for (int copy = 0; copy < 2; ++copy) {
                    const bool vis = copy ? (fflags & 8u) : (fflags & 2u);
                    if (copy && !(fflags & 4u)) break;
                    if (!vis && !(fflags & 1u)) continue;
#pragma unroll
                    for (int t = 0; t < 6; ++t) {
                        // the reversed copy swaps edges 0 and 1: its task t has the scan-lines of task t ^ 2 (t < 4)
                        const int ts = (copy && t < 4) ? (t ^ 2) : t;
                        const int n = (ts < 4 ? ns >> (8 * ts) : ns_hi >> (8 * (ts - 4))) & 0xff;
                        for (int k = 0; k < n; ++k, ++idx)
                            if (idx >= lo && idx < lo + B2_CAP)
                                list[idx - lo] = (unsigned)threadIdx.x | ((unsigned)t << 8) | ((unsigned)copy << 11) |
                                                 ((fflags & 1u) << 12) | (vis ? 1u << 13 : 0u) | ((unsigned)k << 14);
                    }
                }
            }
            __syncthreads();
            // ---- (2) a warp takes 32 segments. (a) thread per segment: ownership of the in-pixels (loads issued
            //      together), which crossings are worth a sweep; (b) those crossings are compacted over the lanes,
            //      thread per crossing: sweep range, runs inside it -> (crossing, run) items in the warp's queue;
            //      (c) 32 items are evaluated at a time, one per lane, straight-line code; results return to the
            //      segment's accumulators in shared memory.
            const int n_here = min(B2_CAP, total - lo);
            TaskStash &st = stash[warp];
            SweepQueue2 &sq = sq2[warp];
            int qhead = 0, qtail = 0;
            unsigned short *pr = pairs[warp];
            for (int j0 = warp * 32; j0 < n_here; j0 += NTHREADS) {
                const int j = j0 + lane;
                const bool valid = j < n_here;
                const unsigned ent = valid ? list[j] : 0u;
                const int t = (ent >> 8) & 7, e = t >> 1, axis = t & 1, k = ent >> 14;
                const bool bnd = (ent >> 12) & 1u, vis = (ent >> 13) & 1u;
                BwdFace bf = load_bwd_face(brecs + (valid ? round * fpr + (int)(ent & 0xffu) : 0));
                if ((ent >> 11) & 1u) reverse_bwd_face(bf, F);
                const TaskGeom g = task_geom(bf, e, axis, is);
                const int d0a = g.d0_from + k * B2_SEG;
                const int d0b = valid ? min(g.d0_to, d0a + B2_SEG - 1) : d0a - 1;
                const int lN = axis == 0 ? 2 : 0;
                st.slope[lane] = g.slope; st.p0d0[lane] = g.p0d0; st.p0d1[lane] = g.p0d1; st.p1d0[lane] = g.p1d0;
                st.p2d0[lane] = g.p2d0; st.p2d1[lane] = g.p2d1; st.s02[lane] = g.s02; st.s21[lane] = g.s21;
                st.misc[lane] = d0a | (axis << 12) | (g.dir > 0 ? 1 << 13 : 0);
                sacc[warp][lane][0] = 0.f;
                sacc[warp][lane][1] = 0.f;
                // (a) bit q of `out`: the face owns the in-pixel of scan-line d0a + q and the line has missing coverage
                //     beyond the out-pixel; bit q of `in`: a crossing of a face that can have in-sweeps
                unsigned out = 0, in = 0;
#pragma unroll
                for (int half = 0; half < B2_SEG / 4; ++half) {   // four scan-lines at a time; short segments stop early
                    if (half > 0 && !__any_sync(FULL, d0b >= d0a + 4 * half)) break;
                    int own[4];
                    uint32_t inf[4];
                    int d1o[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int q = 4 * half + q4, d0 = d0a + q;
                        own[q4] = -3; inf[q4] = 0u; d1o[q4] = 0;
                        if (d0 <= d0b) {
                            const float x = g.slope * ((float)d0 - g.p0d0) + g.p0d1;
                            const int d1_in = __float2int_rz(g.dir > 0 ? floorf(x) : ceilf(x));
                            const int d1_out = d1_in + g.dir;
                            if ((unsigned)d1_in < (unsigned)is && (unsigned)d1_out < (unsigned)is) {
                                if (bnd) in |= 1u << q;
                                if (vis) {
                                    own[q4] = __ldg(face_index + (axis == 0 ? d1_in * is + d0 : d0 * is + d1_in));
                                    inf[q4] = __ldg(S.run_info + lN * is + d0);
                                    d1o[q4] = d1_out;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        // extent of the line's missing-coverage list against the sweep [d1_out, border]
                        const int lo_ = (inf[q4] >> 4) & 0xfffu, hi_ = inf[q4] >> 16;
                        const bool beyond = g.dir > 0 ? hi_ >= d1o[q4] : lo_ <= d1o[q4];
                        if (own[q4] == bf.fn && (inf[q4] & 15u) && beyond) out |= 1u << (4 * half + q4);
                    }
                }
                // (b) compaction: pair = lane | q << 5 | in-sweep << 8
                const int mine = __popc(out) + __popc(in);
                int pre = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, pre, o);
                    if (lane >= o) pre += v;
                }
                const int npairs = __shfl_sync(FULL, pre, 31);
                if (npairs == 0) continue;   // (warp-uniform)
                {
                    int at = pre - mine;
                    unsigned m = out;
                    while (m) { const int q = __ffs(m) - 1; m &= m - 1; pr[at++] = (unsigned short)(lane | (q << 5)); }
                    m = in;
                    while (m) { const int q = __ffs(m) - 1; m &= m - 1; pr[at++] = (unsigned short)(lane | (q << 5) | 256); }
                }
                __syncwarp();
                for (int i0 = 0; i0 < npairs; i0 += 32) {
                    unsigned flags = 0;
                    int ra = 0, rc = 0, ls = 0, d0 = 0, owner = 0, r0 = 0, nr = 0;
                    float x = 0.f, K0 = 0.f, K1 = 0.f, wa0 = 0.f, wa1 = 0.f;
                    if (i0 + lane < npairs) {
                        const unsigned pw = pr[i0 + lane];
                        owner = pw & 31;
                        const int q = (pw >> 5) & 7;
                        const int misc = st.misc[owner];
                        const int ax = (misc >> 12) & 1, dir = (misc >> 13) & 1 ? 1 : -1, lNo = ax == 0 ? 2 : 0;
                        d0 = (misc & 0xfff) + q;
                        const float fd0 = (float)d0, p0d0 = st.p0d0[owner], p0d1 = st.p0d1[owner], p1d0 = st.p1d0[owner];
                        x = st.slope[owner] * (fd0 - p0d0) + p0d1;
                        const int d1_in = __float2int_rz(dir > 0 ? floorf(x) : ceilf(x));
                        const int d1_out = d1_in + dir;
                        flags = (p1d0 != fd0 ? 1u : 0u) | (p0d0 != fd0 ? 2u : 0u);
                        bool walk = false;
                        uint32_t info = 0;
                        if (!(pw & 256u)) {   // out-sweep: from the out-pixel to the image border
                            const int lim = dir > 0 ? is - 1 : 0;
                            ra = min(d1_out, lim); rc = max(d1_out, lim);
                            ls = lNo;
                            info = __ldg(S.run_info + ls * is + d0);
                        } else {              // in-sweep: from the in-pixel to the opposite edge of the triangle
                            const uint32_t *cov = ax == 0 ? cov_col : cov_row;   // coverage of line d0 along d1
                            const bool alpha_out = (__ldg(cov + d0 * W + (d1_out >> 5)) >> (d1_out & 31)) & 1u;
                            ls = alpha_out ? lNo : lNo + 1;
                            info = __ldg(S.run_info + ls * is + d0);
                            if ((info & 15u) != 0u) {
                                const float p2d0 = st.p2d0[owner];
                                float c2;
                                if ((fd0 - p0d0) * (fd0 - p2d0) < 0.f) c2 = st.s02[owner] * (fd0 - p0d0) + p0d1;
                                else c2 = st.s21[owner] * (fd0 - p2d0) + st.p2d1[owner];
                                const int lim = __float2int_rz(dir > 0 ? ceilf(c2) : floorf(c2));
                                ra = max(min(d1_in, lim), 0); rc = min(max(d1_in, lim), is - 1);
                                // an in-sweep can straddle its crossing: by a pixel when the triangle is thinner than a
                                // pixel there (eval_item copes with NEAR_N - 1 pixels on the near side), by many when its
                                // far end is extrapolated (irregular faces): those walk the bit line
                                walk = (float)ra < x - (float)(NEAR_N - 1) && (float)rc > x;
                                if (ra > rc) info = 0u;
                            }
                        }
                        // runs of the line's list with pixels inside [ra, rc]: the lists are sorted, so they are
                        // neighbours [r0, r0 + nr)
                        const unsigned cnt = info & 15u;
                        if (cnt != 0u && rc >= (int)((info >> 4) & 0xfffu) && ra <= (int)(info >> 16)) {
                            const uint2 *rl = S.runs + (ls * is + d0) * RCAP;
                            if (cnt == RUN_OVERFLOW || walk) {
                                nr = -1;
                            } else {
                                for (unsigned r = 0; r < cnt; ++r) {
                                    const unsigned se = __ldg(&rl[r].x);
                                    if (rc >= (int)(se & 0xffffu) && ra <= (int)(se >> 16)) {
                                        if (nr == 0) r0 = r;
                                        ++nr;
                                    }
                                }
                            }
                        }
                        if (nr) {
                            const float ka = p1d0 - p0d0;
                            K0 = __fdividef(ka, p1d0 - fd0) * S.inv_is2;
                            K1 = __fdividef(ka, fd0 - p0d0) * S.inv_is2;
                        }
                        if (nr < 0) {   // rare: run list overflowed or straddling sweep, walk the bit line in place
                            sweep_walk(S, ls, d0, ra, rc, x, K0, K1, flags, wa0, wa1);
                            nr = 0;
                        }
                    }
                    if (__any_sync(FULL, wa0 != 0.f || wa1 != 0.f)) {   // walked sums: lanes in order (fixed summation order)
                        for (int l = 0; l < 32; ++l) {
                            const float v0 = __shfl_sync(FULL, wa0, l), v1 = __shfl_sync(FULL, wa1, l);
                            const int ow = __shfl_sync(FULL, owner, l);
                            if (lane == 0 && (v0 != 0.f || v1 != 0.f)) { sacc[warp][ow][0] += v0; sacc[warp][ow][1] += v1; }
                        }
                        __syncwarp();
                    }
                    // (c) sweeps with runs go to the warp's queue in pair order (the sweeps of one segment stay
                    //     neighbours); 32 are evaluated at a time
                    {
                        const unsigned m = __ballot_sync(FULL, nr > 0);
                        if (nr > 0) {
                            const int at = (qtail + __popc(m & ((1u << lane) - 1u))) & (B2_QCAP - 1);
                            sq.x[at] = x; sq.K0[at] = K0; sq.K1[at] = K1;
                            sq.meta[at] = flags | ((unsigned)owner << 2) | ((unsigned)r0 << 7) | ((unsigned)nr << 10);
                            sq.where[at] = (unsigned)ra | ((unsigned)rc << 10) | ((unsigned)(ls * is + d0) << 20);
                        }
                        qtail += __popc(m);
                        __syncwarp();
                        if (qtail - qhead >= 32) { b2_drain(sq, qhead, 32, S.runs, S.eps, sacc[warp]); qhead += 32; }
                    }
                }
                if (qtail > qhead) { b2_drain(sq, qhead, qtail - qhead, S.runs, S.eps, sacc[warp]); qhead = qtail; }
                __syncwarp();
                // slot pi0*3 + (1 - axis): axis 0 sweeps along y and yields the y gradient, axis 1 the x gradient
                const float acc0 = sacc[warp][lane][0], acc1 = sacc[warp][lane][1];
                if (acc0 != 0.f) hm_accumulate(grad_ndc, grad_fixed, g.vid0 * 3 + (1 - axis), acc0);
                if (acc1 != 0.f) hm_accumulate(grad_ndc, grad_fixed, g.vid1 * 3 + (1 - axis), acc1);
                __syncwarp();
            }
        }
    }
}

// @region after
// ------------------------------------------------------------------------------------------ RGB / depth (visualisation)
// Forward of nr.rasterize_rgbad for texture_size 1 on top of the face_index map: a covered raster sample takes the
// lit colour of its face (doubled numbering: f or F + f) and the reference's interpolated depth (same arithmetic
// as the z-buffer of the oracle), an uncovered one the background colour and `far`; vertical flip and, with
// anti-aliasing, the 2x2 average. One thread per output pixel. Visualisation only (no backward).
__global__ void __launch_bounds__(NTHREADS)
raster_shade_kernel(const FaceRec *__restrict__ recs, const int32_t *__restrict__ face_index,
                    const float *__restrict__ colours, int colours_batch, int B, int F, int is, int aa, float far_,
                    float bg0, float bg1, float bg2, float *__restrict__ rgb, float *__restrict__ depth) {
    const int R = aa ? is / 2 : is;
    const long i = (long)blockIdx.x * NTHREADS + threadIdx.x;
    if (i >= (long)B * R * R) return;
    const int b = (int)(i / ((long)R * R)), r = (int)((i / R) % R), c = (int)(i % R);
    const int n = aa ? 2 : 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // torch's avg_pool2d sums the window row by row (flipped rows 2r, 2r + 1 = raster rows is-1-2r, is-2-2r)
    for (int dy = 0; dy < n; ++dy)
        for (int dx = 0; dx < n; ++dx) {
            const int yi = is - 1 - (n * r + dy), xi = n * c + dx;
            const int fn = face_index[((long)b * is + yi) * is + xi];
            float px[4] = {bg0, bg1, bg2, far_};
            if (fn >= 0) {
                const FaceRec &rec = recs[(long)b * F + (fn >= F ? fn - F : fn)];
                const float *col = colours + ((long)(colours_batch > 1 ? b : 0) * 2 * F + fn) * 3;
                px[0] = col[0]; px[1] = col[1]; px[2] = col[2];
                px[3] = exact_depth(rec.inv, xi, yi, rec.c[2], rec.c[5], rec.c[8]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += px[k];
        }
    const float s = aa ? 0.25f : 1.f;
    const long plane = (long)R * R, o = (long)r * R + c;
    if (rgb) {
        rgb[((long)b * 3 + 0) * plane + o] = acc[0] * s;
        rgb[((long)b * 3 + 1) * plane + o] = acc[1] * s;
        rgb[((long)b * 3 + 2) * plane + o] = acc[2] * s;
    }
    if (depth) depth[(long)b * plane + o] = acc[3] * s;
}

// nr.lighting for flat per-face colours (texture_size 1), both windings of every face: normal of the 3-D triangle
// n = (v0 - v1) x (v2 - v1) / max(|.|, 1e-5), light = ambient + directional * relu(n . direction), lit = colour * light.
// lit [B, 2F, 3]: face f, then its reversed (fill_back) copy F + f (normal negated; zeros when fill_back is off).
__global__ void __launch_bounds__(NTHREADS)
face_lighting_kernel(const float *__restrict__ verts, const int32_t *__restrict__ faces, int faces_batch,
                     const float *__restrict__ colours, int colours_batch, int B, int V, int F, int fill_back, float ia,
                     float id, float ca0, float ca1, float ca2, float cd0, float cd1, float cd2, float d0, float d1,
                     float d2, float *__restrict__ lit) {
    const long i = (long)blockIdx.x * NTHREADS + threadIdx.x;
    if (i >= (long)B * F) return;
    const int b = (int)(i / F), f = (int)(i % F);
    const int32_t *fc = faces + ((faces_batch > 1 ? (long)b * F : 0) + f) * 3;
    const float *col = colours + ((colours_batch > 1 ? (long)b * F : 0) + f) * 3;
    float p[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int vi = min(max(fc[k], 0), V - 1);
        const float *q = verts + ((long)b * V + vi) * 3;
        p[k][0] = q[0]; p[k][1] = q[1]; p[k][2] = q[2];
    }
    const float a[3] = {p[0][0] - p[1][0], p[0][1] - p[1][1], p[0][2] - p[1][2]};
    const float c[3] = {p[2][0] - p[1][0], p[2][1] - p[1][1], p[2][2] - p[1][2]};
    float n[3] = {a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]};
    const float inv = 1.f / fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-5f);
    const float cs = (n[0] * d0 + n[1] * d1 + n[2] * d2) * inv;
    const float amb[3] = {ia * ca0, ia * ca1, ia * ca2}, dirc[3] = {id * cd0, id * cd1, id * cd2};
    float *o0 = lit + ((long)b * 2 * F + f) * 3, *o1 = lit + ((long)b * 2 * F + F + f) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o0[k] = col[k] * (amb[k] + dirc[k] * fmaxf(cs, 0.f));
        o1[k] = fill_back ? col[k] * (amb[k] + dirc[k] * fmaxf(-cs, 0.f)) : 0.f;
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

// The per-image pixel boxes follow the face boxes in the bbox buffer (HM_FACE_BBOX_BUFFER_BYTES), 16-byte aligned.
inline int *image_boxes(void *bboxes, int B, int F) {
    const size_t off = ((size_t)B * F * HM_FACE_BBOX_BYTES + 15) & ~(size_t)15;
    return reinterpret_cast<int *>(static_cast<char *>(bboxes) + off);
}

// ------------------------------------------------------------------------------------------ fused loss + sweep lists
// hm_sil_loss_fwd_bwd + hm_raster_grad_prep in one kernel, CTA per image. The silhouette loss gradient takes nine
// values per image (gscale * k / 4, k = -4 .. 4: alpha is a multiple of 1/4, ref and keep are 0 / 1), so the whole
// gradient map stays in shared memory as one signed byte per output pixel while the kernel (A) streams alpha and
// the target once (loss, IoU, grad_alpha) and (B) derives, thread per raster line, the four sweep bit lines and
// their run-length form from those bytes and the coverage words - the gradient is never re-read from memory (the
// unfused pair re-reads it twice and gathers it once more per set bit). Same outputs, bit for bit.
constexpr int SLP_THREADS = 512;
template <bool AA>
__global__ void __launch_bounds__(SLP_THREADS, 2)
sil_loss_prep_kernel(const float *__restrict__ alpha, const int8_t *__restrict__ target, const float *__restrict__ norm,
                     float weight, int R, float *__restrict__ loss_img, int loss_stride, float *__restrict__ iou_img,
                     int iou_stride, float *__restrict__ grad_alpha, const uint32_t *__restrict__ cov_row,
                     const uint32_t *__restrict__ cov_col, uint32_t *__restrict__ m_row, uint32_t *__restrict__ m_col,
                     uint2 *__restrict__ runs, uint32_t *__restrict__ run_counts) {
    extern __shared__ __align__(16) signed char code[];   // [R][R + 4]: 4 * keep * (keep * alpha - ref) * sign(gscale)
    __shared__ float scratch[3 * 32];
    const int b = blockIdx.x, npix = R * R;
    const int CS = R + 4;                                  // row stride of the byte map
    unsigned *sgn_row = reinterpret_cast<unsigned *>(code + (size_t)R * CS);   // [2][R][R / 32]: byte < 0, byte > 0
    unsigned *sgn_col = sgn_row + 2 * R * (R / 32);                            // the same by column (bit = row)
    const int is = AA ? 2 * R : R, W = is / 32;
    const float nb = norm[b];
    const float gscale = 2.f * weight * nb;
    const int sgn = gscale > 0.f ? 1 : (gscale < 0.f ? -1 : 0);
    // ---- (A) loss, IoU, gradient
    {
        const float4 *a4 = reinterpret_cast<const float4 *>(alpha + (long)b * npix);
        const char4 *t4 = reinterpret_cast<const char4 *>(target + (long)b * npix);
        float4 *g4 = reinterpret_cast<float4 *>(grad_alpha + (long)b * npix);
        float acc[3] = {0.f, 0.f, 0.f};  // sum sq, intersection, union
        for (int i = threadIdx.x; i < npix / 4; i += SLP_THREADS) {
            const float4 a = a4[i];
            const char4 t = t4[i];
            const float av[4] = {a.x, a.y, a.z, a.w};
            const int tv[4] = {t.x, t.y, t.z, t.w};
            float gv[4];
            char4 c;
            signed char *cv = reinterpret_cast<signed char *>(&c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float keep = tv[k] >= 0 ? 1.f : 0.f, ref = tv[k] > 0 ? 1.f : 0.f;
                const float img = keep * av[k];
                const float d = img - ref;
                acc[0] += d * d;
                acc[1] += img * ref;
                acc[2] += fminf(fmaxf(img + ref, 0.f), 1.f);
                gv[k] = gscale * keep * d;
                cv[k] = (signed char)(__float2int_rn(4.f * keep * d) * sgn);
            }
            g4[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
            *reinterpret_cast<char4 *>(code + ((4 * i) / R) * CS + (4 * i) % R) = c;
        }
        __syncthreads();   // (block_sum's own barriers would do; explicit for the code bytes)
        // block sum over SLP_THREADS / 32 warps
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) { scratch[warp] = acc[0]; scratch[32 + warp] = acc[1]; scratch[64 + warp] = acc[2]; }
        __syncthreads();
        if (warp == 0) {
            float x0 = lane < SLP_THREADS / 32 ? scratch[lane] : 0.f, x1 = lane < SLP_THREADS / 32 ? scratch[32 + lane] : 0.f,
                  x2 = lane < SLP_THREADS / 32 ? scratch[64 + lane] : 0.f;
            x0 = warp_sum(x0); x1 = warp_sum(x1); x2 = warp_sum(x2);
            if (lane == 0) {
                if (loss_img) loss_img[(long)b * loss_stride] = x0 * nb;
                if (iou_img) iou_img[(long)b * iou_stride] = x1 / (x2 + 1e-6f);
            }
        }
    }
    // |grad| at raster resolution of a pixel whose byte is +-k (the expression build_runs evaluates on grad_alpha)
    float Gk[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float gval = gscale * 1.f * ((float)k * 0.25f);
        Gk[k] = fabsf(AA ? 0.25f * gval : gval);
    }
    // ---- (B1) sign bit masks of the byte map, by output row (bit = column) and by output column (bit = row): one
    //      ballot per 32 pixels (the row stride of the bytes is R + 4, so a warp reading down a column is conflict-free)
    const int RW = R / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += SLP_THREADS / 32)
    for (int w = 0; w < RW; ++w) {
        const int v = code[r * CS + 32 * w + lane];
        const unsigned ng = __ballot_sync(0xffffffffu, v < 0), ps = __ballot_sync(0xffffffffu, v > 0);
        if (lane == 0) { sgn_row[(0 * R + r) * RW + w] = ng; sgn_row[(1 * R + r) * RW + w] = ps; }
        const int vc = code[(32 * w + lane) * CS + r];   // column r, rows 32 w ..
        const unsigned ngc = __ballot_sync(0xffffffffu, vc < 0), psc = __ballot_sync(0xffffffffu, vc > 0);
        if (lane == 0) { sgn_col[(0 * R + r) * RW + w] = ngc; sgn_col[(1 * R + r) * RW + w] = psc; }
    }
    __syncthreads();
    // ---- (B2) thread per raster line: rows (lists 0 / 1), then columns (lists 2 / 3)
    const long plane = (long)is * W;
    for (int ln = threadIdx.x; ln < 2 * is; ln += SLP_THREADS) {
        const bool col = ln >= is;
        const int line = col ? ln - is : ln;
        const uint32_t *cov = (col ? cov_col : cov_row) + ((long)b * is + line) * W;
        uint32_t *mo = (col ? m_col : m_row) + ((long)b * 2 * is + line) * W;
        // the line of the byte map behind this raster line: row lines read output row `o` left to right, column lines
        // read output column `o` from the bottom row up (raster y grows as the output row shrinks)
        const int o = col ? (AA ? line >> 1 : line) : (AA ? (is - 1 - line) >> 1 : is - 1 - line);
        const unsigned *sg = (col ? sgn_col : sgn_row) + o * RW;
        int n[2] = {0, 0}, rs[2] = {-1, -1}, re[2] = {-1, -1}, first[2] = {0, 0}, rk[2] = {0, 0};
        uint2 *out[2] = {runs + (((long)b * 4 + (col ? 2 : 0)) * is + line) * RCAP,
                         runs + (((long)b * 4 + (col ? 3 : 1)) * is + line) * RCAP};
        // (coverage words of the line are fetched four at a time: one dependent global load per word would serialise it)
        uint4 cov4 = make_uint4(0u, 0u, 0u, 0u);
        for (int w = 0; w < W; ++w) {
            if ((w & 3) == 0) cov4 = __ldg(reinterpret_cast<const uint4 *>(cov) + (w >> 2));
            unsigned neg, pos;
            if (AA) {   // 16 output pixels per word
                const int p0 = col ? R - 16 * (w + 1) : 16 * w;   // first output pixel (lowest index) of the word
                neg = (sg[p0 >> 5] >> (p0 & 31)) & 0xffffu;
                pos = (sg[R * RW + (p0 >> 5)] >> (p0 & 31)) & 0xffffu;
                if (col) { neg = __brev(neg) >> 16; pos = __brev(pos) >> 16; }
                neg = dup_bits16(neg); pos = dup_bits16(pos);
            } else {
                const int wi = col ? RW - 1 - w : w;
                neg = sg[wi]; pos = sg[R * RW + wi];
                if (col) { neg = __brev(neg); pos = __brev(pos); }
            }
            const unsigned A = (w & 3) == 0 ? cov4.x : (w & 3) == 1 ? cov4.y : (w & 3) == 2 ? cov4.z : cov4.w;
            const unsigned m[2] = {neg & ~A, pos & A};
            mo[w] = m[0];
            mo[w + plane] = m[1];
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                unsigned bits = m[l];
                while (bits) {
                    // next stretch of consecutive set bits [lo, hi] of the word, then its pixels grouped by |byte|
                    const int lo = __ffs(bits) - 1;
                    const unsigned rest = ~(bits >> lo);
                    const int len = rest ? __ffs(rest) - 1 : 32 - lo;
                    const int hi = lo + len - 1;
                    bits = hi >= 31 ? 0u : bits & (0xffffffffu << (hi + 1));
                    int p = lo;
                    while (p <= hi) {
                        const int d1 = w * 32 + p;
                        const int op = AA ? d1 >> 1 : d1;   // output pixel along the line (raster order)
                        int k = col ? code[(R - 1 - op) * CS + o] : code[o * CS + op];
                        k = k < 0 ? -k : k;
                        const int q = AA ? min(hi, p | 1) : p;   // last raster pixel of this output pixel in the stretch
                        if (rs[l] < 0) first[l] = d1;
                        if (rs[l] >= 0 && d1 == re[l] + 1 && k == rk[l]) {
                            re[l] = w * 32 + q;
                        } else {
                            if (rs[l] >= 0) {
                                if (n[l] < RCAP) out[l][n[l]] = make_uint2((unsigned)rs[l] | ((unsigned)re[l] << 16), __float_as_uint(Gk[rk[l]]));
                                ++n[l];
                            }
                            rs[l] = d1;
                            re[l] = w * 32 + q;
                            rk[l] = k;
                        }
                        p = q + 1;
                    }
                }
            }
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (rs[l] >= 0) {
                if (n[l] < RCAP) out[l][n[l]] = make_uint2((unsigned)rs[l] | ((unsigned)re[l] << 16), __float_as_uint(Gk[rk[l]]));
                ++n[l];
            }
            run_counts[((long)b * 4 + (col ? 2 : 0) + l) * is + line] =
                (n[l] > RCAP ? RUN_OVERFLOW : (unsigned)n[l]) | ((unsigned)first[l] << 4) | ((unsigned)max(re[l], 0) << 16);
        }
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
    HM_NVTX("hm_project_fwd");
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
    HM_NVTX("hm_project_bwd");
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
    HM_NVTX("hm_raster_setup");
    HM_REQUIRE(B >= 0 && V >= 0 && F >= 0 && (faces_batch == 1 || faces_batch == B), "hm_raster_setup: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if ((long)B * F == 0) return HM_OK;  // nothing to set up (empty batch or empty mesh)
    HM_REQUIRE(ndc && faces && records && bboxes, "hm_raster_setup: null pointer");
    HM_REQUIRE(V > 0, "hm_raster_setup: bad sizes");
    const long n = (long)B * F;
    HM_UNSUPPORTED(2L * F > BWD_FN_MASK, "hm_raster_setup: too many faces (%d)", F);
    int *img_box = image_boxes(bboxes, B, F);
    {   // preset of the per-image pixel boxes (a memset node when the stream is being captured)
        const cudaError_t e = cudaMemsetAsync(img_box, 0x7f, (size_t)B * 16, hm_stream(stream));
        if (e != cudaSuccess) {
            hm_set_error("hm_raster_setup: cudaMemsetAsync: %s", cudaGetErrorString(e));
            return HM_ERR_CUDA;
        }
    }
    face_setup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, hm_stream(stream)>>>(
        ndc, faces, faces_batch, B, V, F, is, fill_back, static_cast<FaceRec *>(records),
        reinterpret_cast<BwdRec *>(static_cast<FaceRec *>(records) + n), static_cast<FaceBox *>(bboxes), img_box);
    HM_CHECK_LAUNCH("hm_raster_setup");
    return HM_OK;
}

int hm_raster_sil_fwd(const void *records, const void *bboxes, int B, int F, int image_size,
                      int anti_aliasing, float near_, float far_, int32_t *face_index, float *alpha,
                      uint32_t *cov_row, uint32_t *cov_col, uint32_t *face_vis, uint8_t *cov_blocks, void *stream) {
    HM_NVTX("hm_raster_sil_fwd");
    HM_REQUIRE(B >= 0 && F >= 0 && B <= 65535, "hm_raster_sil_fwd: bad sizes (B <= 65535)");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0) return HM_OK;
    HM_REQUIRE((F == 0 || (records && bboxes)) && face_index && alpha, "hm_raster_sil_fwd: null pointer");
    dim3 grid((is / TILE) * (is / TILE), B);
    const int fwd_smem = TILE * TILE * (int)sizeof(unsigned long long);
    static HmSmemOptIn opt_in;  // static + dynamic shared memory exceeds the 48 KB default
    if (int rc = hm_smem_opt_in(raster_fwd_kernel, fwd_smem, opt_in, "hm_raster_sil_fwd")) return rc;
    if (face_vis && F > 0) {   // the tiles OR into it (a memset node when the stream is being captured)
        const cudaError_t e = cudaMemsetAsync(face_vis, 0, (size_t)B * HM_FACE_VIS_WORDS(F) * sizeof(uint32_t), hm_stream(stream));
        if (e != cudaSuccess) {
            hm_set_error("hm_raster_sil_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
            return HM_ERR_CUDA;
        }
    }
    raster_fwd_kernel<<<grid, NTHREADS, fwd_smem, hm_stream(stream)>>>(
        static_cast<const FaceRec *>(records), static_cast<const FaceBox *>(bboxes), F, is, anti_aliasing, near_, far_,
        face_index, alpha, cov_row, cov_col, face_vis, cov_blocks, image_boxes(const_cast<void *>(bboxes), B, F));
    HM_CHECK_LAUNCH("hm_raster_sil_fwd");
    return HM_OK;
}

int hm_raster_grad_prep(const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, int B,
                        int image_size, int anti_aliasing, uint32_t *m_row, uint32_t *m_col, void *runs,
                        uint32_t *run_counts, void *stream) {
    HM_NVTX("hm_raster_grad_prep");
    HM_REQUIRE(B >= 0 && B <= 65535, "hm_raster_grad_prep: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0) return HM_OK;
    HM_REQUIRE(grad_alpha && cov_row && cov_col && m_row && m_col && runs && run_counts,
               "hm_raster_grad_prep: null pointer");
    const int out_size = anti_aliasing ? is / 2 : is;
    HM_UNSUPPORTED(out_size % 32 != 0, "hm_raster_grad_prep: image size %d must be a multiple of 32", out_size);
    dim3 grid((2 * out_size * (is / 32) + NTHREADS - 1) / NTHREADS, B);
    if (anti_aliasing)
        grad_prep_kernel<true><<<grid, NTHREADS, 0, hm_stream(stream)>>>(grad_alpha, cov_row, cov_col, is, m_row, m_col);
    else
        grad_prep_kernel<false><<<grid, NTHREADS, 0, hm_stream(stream)>>>(grad_alpha, cov_row, cov_col, is, m_row, m_col);
    HM_CHECK_LAUNCH("hm_raster_grad_prep");
    dim3 grid2((4 * is + NTHREADS - 1) / NTHREADS, B);
    build_runs_kernel<<<grid2, NTHREADS, 0, hm_stream(stream)>>>(grad_alpha, m_row, m_col, is, anti_aliasing,
                                                                 static_cast<uint2 *>(runs), run_counts);
    HM_CHECK_LAUNCH("hm_raster_grad_prep(runs)");
    return HM_OK;
}

int hm_raster_sil_bwd(const void *records, const void *bboxes, const int32_t *face_index,
                      const float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col,
                      const uint32_t *face_vis, const uint8_t *cov_blocks,
                      const uint32_t *m_row, const uint32_t *m_col, const void *runs, const uint32_t *run_counts,
                      int B, int V, int F, int image_size, int anti_aliasing, float eps, float *grad_ndc,
                      unsigned long long *grad_fixed, void *stream) {
    HM_NVTX("hm_raster_sil_bwd");
    HM_REQUIRE(B >= 0 && F >= 0 && V >= 0 && B <= 65535, "hm_raster_sil_bwd: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0 || F == 0) return HM_OK;
    HM_REQUIRE(records && bboxes && face_index && grad_alpha && cov_row && cov_col && face_vis && cov_blocks && m_row &&
                   m_col && runs && run_counts && grad_ndc,
               "hm_raster_sil_bwd: null pointer");
    HM_REQUIRE(V > 0, "hm_raster_sil_bwd: bad sizes");
    HM_UNSUPPORTED(is > 1024, "hm_raster_sil_bwd: raster size %d > 1024 is not supported", is);
    const BwdRec *brecs = reinterpret_cast<const BwdRec *>(static_cast<const FaceRec *>(records) + (long)B * F);
    const int fpr = F <= HM_BWD_SMALL_MESH ? NTHREADS / 2 : NTHREADS;   // faces per CTA round (see the kernel)
    const int n_rounds = (F + fpr - 1) / fpr;
    raster_bwd_kernel<<<dim3(min(n_rounds, 8), B), NTHREADS, 0, hm_stream(stream)>>>(
        brecs, static_cast<const FaceBox *>(bboxes), F, V, is, anti_aliasing, eps, face_index, grad_alpha, cov_row,
        cov_col, face_vis, cov_blocks, m_row, m_col, static_cast<const uint2 *>(runs), run_counts, grad_ndc, grad_fixed,
        fpr);
    HM_CHECK_LAUNCH("hm_raster_sil_bwd");
    return HM_OK;
}

int hm_sil_loss_fwd_bwd(const float *alpha, const int8_t *target, const float *norm, float weight, int B,
                        int image_size, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                        float *grad_alpha, void *stream) {
    HM_NVTX("hm_sil_loss_fwd_bwd");
    HM_REQUIRE(alpha && target && norm, "hm_sil_loss_fwd_bwd: null pointer");
    HM_REQUIRE(B >= 0 && image_size > 0 && (image_size * image_size) % 4 == 0, "hm_sil_loss_fwd_bwd: bad sizes");
    if (B == 0) return HM_OK;
    sil_loss_kernel<<<B, NTHREADS, 0, hm_stream(stream)>>>(alpha, target, norm, weight, image_size * image_size,
                                                           loss_img, loss_stride, iou_img, iou_stride, grad_alpha);
    HM_CHECK_LAUNCH("hm_sil_loss_fwd_bwd");
    return HM_OK;
}

int hm_face_lighting(const float *verts, const int32_t *faces, int faces_batch, const float *colours, int colours_batch,
                     int B, int V, int F, int fill_back, float intensity_ambient, float intensity_directional,
                     const float *color_ambient, const float *color_directional, const float *direction, float *lit,
                     void *stream) {
    HM_NVTX("hm_face_lighting");
    HM_REQUIRE(B >= 0 && V >= 0 && F >= 0 && (faces_batch == 1 || faces_batch == B) && (colours_batch == 1 || colours_batch == B),
               "hm_face_lighting: bad sizes");
    if ((long)B * F == 0) return HM_OK;
    HM_REQUIRE(verts && faces && colours && color_ambient && color_directional && direction && lit && V > 0,
               "hm_face_lighting: null pointer");
    const long n = (long)B * F;   // (the three host vectors are read here, at enqueue time)
    face_lighting_kernel<<<(unsigned)((n + NTHREADS - 1) / NTHREADS), NTHREADS, 0, hm_stream(stream)>>>(
        verts, faces, faces_batch, colours, colours_batch, B, V, F, fill_back, intensity_ambient, intensity_directional,
        color_ambient[0], color_ambient[1], color_ambient[2], color_directional[0], color_directional[1],
        color_directional[2], direction[0], direction[1], direction[2], lit);
    HM_CHECK_LAUNCH("hm_face_lighting");
    return HM_OK;
}

int hm_sil_loss_prep(const float *alpha, const int8_t *target, const float *norm, float weight, int B, int image_size,
                     int anti_aliasing, float *loss_img, int loss_stride, float *iou_img, int iou_stride,
                     float *grad_alpha, const uint32_t *cov_row, const uint32_t *cov_col, uint32_t *m_row,
                     uint32_t *m_col, void *runs, uint32_t *run_counts, void *stream) {
    HM_NVTX("hm_sil_loss_prep");
    HM_REQUIRE(B >= 0 && image_size > 0, "hm_sil_loss_prep: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    HM_UNSUPPORTED(image_size % 32 != 0 || image_size > 384,
                   "hm_sil_loss_prep: image size %d (a multiple of 32, at most 384)", image_size);
    if (B == 0) return HM_OK;
    HM_REQUIRE(alpha && target && norm && grad_alpha && cov_row && cov_col && m_row && m_col && runs && run_counts,
               "hm_sil_loss_prep: null pointer");
    const size_t smem = (size_t)image_size * (image_size + 4) + 4 * (size_t)image_size * (image_size / 32) * sizeof(unsigned);
    static HmSmemOptIn opt_in[2];
    if (anti_aliasing) {
        if (int rc = hm_smem_opt_in(sil_loss_prep_kernel<true>, smem, opt_in[1], "hm_sil_loss_prep")) return rc;
        sil_loss_prep_kernel<true><<<B, SLP_THREADS, smem, hm_stream(stream)>>>(
            alpha, target, norm, weight, image_size, loss_img, loss_stride, iou_img, iou_stride, grad_alpha, cov_row,
            cov_col, m_row, m_col, static_cast<uint2 *>(runs), run_counts);
    } else {
        if (int rc = hm_smem_opt_in(sil_loss_prep_kernel<false>, smem, opt_in[0], "hm_sil_loss_prep")) return rc;
        sil_loss_prep_kernel<false><<<B, SLP_THREADS, smem, hm_stream(stream)>>>(
            alpha, target, norm, weight, image_size, loss_img, loss_stride, iou_img, iou_stride, grad_alpha, cov_row,
            cov_col, m_row, m_col, static_cast<uint2 *>(runs), run_counts);
    }
    HM_CHECK_LAUNCH("hm_sil_loss_prep");
    return HM_OK;
}

int hm_raster_shade(const void *records, const int32_t *face_index, const float *colours, int colours_batch, int B,
                    int F, int image_size, int anti_aliasing, float far_, float bg_r, float bg_g, float bg_b,
                    float *rgb, float *depth, void *stream) {
    HM_NVTX("hm_raster_shade");
    HM_REQUIRE(B >= 0 && F >= 0 && (colours_batch == 1 || colours_batch == B), "hm_raster_shade: bad sizes");
    int is;
    if (int rc = check_raster_size(image_size, anti_aliasing, &is)) return rc;
    if (B == 0) return HM_OK;
    HM_REQUIRE(face_index && (F == 0 || (records && colours)) && (rgb || depth), "hm_raster_shade: null pointer");
    const long n = (long)B * image_size * image_size;
    raster_shade_kernel<<<(unsigned)((n + NTHREADS - 1) / NTHREADS), NTHREADS, 0, hm_stream(stream)>>>(
        static_cast<const FaceRec *>(records), face_index, colours, colours_batch, B, F, is, anti_aliasing, far_, bg_r,
        bg_g, bg_b, rgb, depth);
    HM_CHECK_LAUNCH("hm_raster_shade");
    return HM_OK;
}

}  // extern "C"
