// common.cuh -- shared device helpers for the sm_100a aggregation kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <climits>

namespace gnnagg {

// Edges of the CSR (or of a scheduled group list) are cut into fixed-size "items"; the row that
// contains the first edge of every kFineItem-edge block is precomputed once per graph
// (item_row[]), so any kernel can start walking rows at an arbitrary edge position.
constexpr int kFineItem = 128;
// edges staged per warp in the aggregation kernels (= items of 512 / (32/lanes_per_row) edges)
constexpr int kWarpEdges = 512;
constexpr int kCtaWarps = 8;
constexpr int kCtaThreads = kCtaWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D bulk (TMA) copy global -> shared ------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 128-bit shared loads from a 32-bit shared address held in a register (see opaque32)
__device__ __forceinline__ int4 lds_i4(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// Makes a loop-invariant value opaque to the optimiser so that it stays in a register instead of being
// re-derived from %tid / kernel parameters in every iteration of the hot loop (ptxas rematerialises address
// arithmetic under register pressure; in the gather loop that cost ~15 instructions per 8 edges).
__device__ __forceinline__ uint32_t opaque32(uint32_t x)
{
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint64_t opaque64(uint64_t x)
{
    uint64_t y;
    asm volatile("mov.u64 %0, %1;" : "=l"(y) : "l"(x));
    return y;
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialisation attribute (capi.cu: launch_dep)
// may be scheduled while its predecessor in the stream drains; it must not touch what the predecessor wrote before this
// returns.  Without the attribute the instruction is a no-op.  Every thread calls it, ahead of any early return, so that
// the completion of the dependent grid always implies the completion of its predecessor.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- vector loads / stores -------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stg_f4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void fma4(float4 &acc, float w, const float4 &v)
{
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}
// acc += max(a + b, 0): the per-edge op of the MLP aggregator after the projection has been hoisted
__device__ __forceinline__ void relu_add4(float4 &acc, const float4 &a, const float4 &b)
{
    acc.x += fmaxf(a.x + b.x, 0.f);
    acc.y += fmaxf(a.y + b.y, 0.f);
    acc.z += fmaxf(a.z + b.z, 0.f);
    acc.w += fmaxf(a.w + b.w, 0.f);
}
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b)
{
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ bool is_zero4(const float4 &a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f; }
// 128-bit reduction into global memory (RED.E.ADD.F32x4 on sm_90+)
__device__ __forceinline__ void red_add_f4(float *p, const float4 &v) { atomicAdd(reinterpret_cast<float4 *>(p), v); }

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b)
{
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Reduce-scatter of CNT per-lane values over a virtual warp of LPR lanes (LPR >= CNT, both powers of two): at
// every butterfly step half of the values travel and half stay, so CNT values cost CNT-1 + log2(LPR/CNT)
// shuffles instead of CNT*log2(LPR).  Afterwards d[0] of lane vl holds the total of value index
// vl / (LPR/CNT); the LPR/CNT lanes of such a group hold the same total.
template <int LPR, int OFF, int CNT>
struct VwReduceScatter {
    static __device__ __forceinline__ void run(float *d, int vl, unsigned mask)
    {
        if constexpr (OFF >= 1) {
            if constexpr (CNT > 1) {
                constexpr int H = CNT / 2;
                const bool upper = (vl & OFF) != 0;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float send = upper ? d[i] : d[i + H];
                    const float keep = upper ? d[i + H] : d[i];
                    d[i] = keep + __shfl_xor_sync(mask, send, OFF, LPR);
                }
                VwReduceScatter<LPR, OFF / 2, H>::run(d, vl, mask);
            } else {
                d[0] += __shfl_xor_sync(mask, d[0], OFF, LPR);
                VwReduceScatter<LPR, OFF / 2, 1>::run(d, vl, mask);
            }
        }
    }
};

// row that contains edge e: last r with ptr[r] <= e.  item_row narrows the search to the rows
// that intersect the kFineItem-edge block of e.
__device__ __forceinline__ int row_of_edge(const int *__restrict__ ptr, const int *__restrict__ item_row,
                                           int num_items, int num_rows, int e)
{
    const int b = e / kFineItem;
    int lo = __ldg(item_row + b);
    int hi = (b + 1 < num_items) ? __ldg(item_row + b + 1) : num_rows - 1;
    // invariant: ptr[lo] <= e, answer in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(ptr + mid) <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

}  // namespace gnnagg
