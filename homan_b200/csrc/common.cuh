// Shared helpers of the homan_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler (nsys, ncu --nvtx) is attached

#include "../../include/homan_b200.h"

void hm_set_error(const char *fmt, ...);

#define HM_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            hm_set_error(__VA_ARGS__);        \
            return HM_ERR_INVALID;            \
        }                                     \
    } while (0)

#define HM_UNSUPPORTED(cond, ...)             \
    do {                                      \
        if (cond) {                           \
            hm_set_error(__VA_ARGS__);        \
            return HM_ERR_UNSUPPORTED;        \
        }                                     \
    } while (0)

// cudaPeekAtLastError is legal during stream capture and does not clear sticky errors.
#define HM_CHECK_LAUNCH(name)                                                         \
    do {                                                                              \
        cudaError_t e__ = cudaPeekAtLastError();                                      \
        if (e__ != cudaSuccess) {                                                     \
            hm_set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));         \
            (void)cudaGetLastError();                                                 \
            return HM_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

static inline cudaStream_t hm_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// One NVTX range per C-ABI call (named after the entry point): the host-side enqueue of the call's launches.
struct HmNvtxRange {
    explicit HmNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~HmNvtxRange() { nvtxRangePop(); }
    HmNvtxRange(const HmNvtxRange &) = delete;
    HmNvtxRange &operator=(const HmNvtxRange &) = delete;
};
#define HM_NVTX(name) HmNvtxRange hm_nvtx_range__(name)

// Opt-in to more than 48 KB of dynamic shared memory. The attribute belongs to the (device, kernel) pair, so the
// bookkeeping is per device: one slot per ordinal, written with relaxed atomics (two host threads racing on the same
// device both issue the call, which is idempotent). `state` is one zero-initialised array per kernel.
constexpr int HM_MAX_DEVICES = 64;
struct HmSmemOptIn {
    std::atomic<size_t> configured[HM_MAX_DEVICES];
};
template <class Kernel>
static inline int hm_smem_opt_in(Kernel kernel, size_t bytes, HmSmemOptIn &state, const char *who) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev >= 0 && dev < HM_MAX_DEVICES &&
        state.configured[dev].load(std::memory_order_relaxed) >= bytes)
        return HM_OK;
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        hm_set_error("%s: cudaFuncSetAttribute(%zu bytes of shared memory): %s", who, bytes, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return HM_ERR_CUDA;
    }
    if (dev >= 0 && dev < HM_MAX_DEVICES) state.configured[dev].store(bytes, std::memory_order_relaxed);
    return HM_OK;
}

// Order-independent accumulation (test mode): contributions are rounded to 2^-44 fixed point and summed with 64-bit
// integer atomics, so the result does not depend on the order in which CTAs and warps arrive (float atomics do).
// |value| must stay below 2^19; hm_fold_fixed converts the sums back.
constexpr float HM_FIXED_SCALE = 17592186044416.f;   // 2^44
__device__ __forceinline__ void hm_accumulate(float *dst, unsigned long long *fixed, long idx, float v) {
    if (fixed) atomicAdd(fixed + idx, (unsigned long long)__float2ll_rn(fminf(fmaxf(v, -5e5f), 5e5f) * HM_FIXED_SCALE));
    else atomicAdd(dst + idx, v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of N values per thread; result valid in thread 0 (and broadcast through `out`).
// `scratch` must hold N * 32 floats. All threads of the block must call it.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float x = lane < nwarps ? scratch[i * 32 + lane] : 0.f;
            x = warp_sum(x);
            if (lane == 0) scratch[i * 32] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = scratch[i * 32];
    __syncthreads();
}

// Block-wide max of N values per thread (same contract as block_sum).
template <int N>
__device__ __forceinline__ void block_max(float (&v)[N], float *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_max(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float x = lane < nwarps ? scratch[i * 32 + lane] : -3.4e38f;
            x = warp_max(x);
            if (lane == 0) scratch[i * 32] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = scratch[i * 32];
    __syncthreads();
}

// rot6d_to_matrix (homan/utils/geometry.py:9-27) and its backward. r6 is the [3,2] row-major parameter;
// R is row-major with columns b1 b2 b3.
struct Rot6d {
    float a1[3], a2[3], b1[3], b2[3], n1, nu, R[9];
};
__device__ inline void rot6d_forward(const float *r6, Rot6d &o) {
    for (int i = 0; i < 3; ++i) { o.a1[i] = r6[2 * i]; o.a2[i] = r6[2 * i + 1]; }
    o.n1 = fmaxf(sqrtf(o.a1[0] * o.a1[0] + o.a1[1] * o.a1[1] + o.a1[2] * o.a1[2]), 1e-12f);
    for (int i = 0; i < 3; ++i) o.b1[i] = o.a1[i] / o.n1;
    const float dp = o.b1[0] * o.a2[0] + o.b1[1] * o.a2[1] + o.b1[2] * o.a2[2];
    float u[3];
    for (int i = 0; i < 3; ++i) u[i] = o.a2[i] - dp * o.b1[i];
    o.nu = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    for (int i = 0; i < 3; ++i) o.b2[i] = u[i] / o.nu;
    const float b3[3] = {o.b1[1] * o.b2[2] - o.b1[2] * o.b2[1], o.b1[2] * o.b2[0] - o.b1[0] * o.b2[2],
                         o.b1[0] * o.b2[1] - o.b1[1] * o.b2[0]};
    for (int i = 0; i < 3; ++i) { o.R[i * 3] = o.b1[i]; o.R[i * 3 + 1] = o.b2[i]; o.R[i * 3 + 2] = b3[i]; }
}
// G = d loss / d R (row-major) -> g6 [3,2] row-major (written, not accumulated)
__device__ inline void rot6d_backward(const Rot6d &o, const float *G, float *g6) {
    float gb1[3] = {G[0], G[3], G[6]}, gb2[3] = {G[1], G[4], G[7]};
    const float gb3[3] = {G[2], G[5], G[8]};
    const float *b1 = o.b1, *b2 = o.b2, *a2 = o.a2;
    gb1[0] += b2[1] * gb3[2] - b2[2] * gb3[1]; gb1[1] += b2[2] * gb3[0] - b2[0] * gb3[2]; gb1[2] += b2[0] * gb3[1] - b2[1] * gb3[0];
    gb2[0] += gb3[1] * b1[2] - gb3[2] * b1[1]; gb2[1] += gb3[2] * b1[0] - gb3[0] * b1[2]; gb2[2] += gb3[0] * b1[1] - gb3[1] * b1[0];
    const float d2 = gb2[0] * b2[0] + gb2[1] * b2[1] + gb2[2] * b2[2];
    float gu[3];
    for (int i = 0; i < 3; ++i) gu[i] = (gb2[i] - d2 * b2[i]) / o.nu;
    const float dp = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float gub1 = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
    float ga2[3];
    for (int i = 0; i < 3; ++i) {
        ga2[i] = gu[i] - gub1 * b1[i];
        gb1[i] += -dp * gu[i] - gub1 * a2[i];
    }
    const float d1 = gb1[0] * b1[0] + gb1[1] * b1[1] + gb1[2] * b1[2];
    for (int i = 0; i < 3; ++i) {
        g6[2 * i] = (gb1[i] - d1 * b1[i]) / o.n1;
        g6[2 * i + 1] = ga2[i];
    }
}

// 1-D bulk asynchronous copy global -> shared through the TMA unit (cp.async.bulk, SASS UBLKCP),
// completion signalled on an mbarrier. Addresses and size must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
