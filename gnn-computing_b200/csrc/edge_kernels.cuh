// edge_kernels.cuh -- edge-parallel kernels: un-fused GAT pieces, edge-wise GCN,
// CSR->edge list, the naive SpMM and the validators of include/spmm.h.
//
// The reference runs all of these warp-per-row (aggr_gat.h:5-92, aggr_sddmm.h:5-83,
// aggregator.h:11-23), which serialises hub rows on one warp.  Here every kernel is parallel
// over EDGES: the row of an edge is recovered with a short binary search bounded by the
// precomputed item_row table (common.cuh: row_of_edge), so work is balanced whatever the degree
// distribution, and row sums use the same item/carry scheme as the aggregation kernels
// (deterministic, no float atomics).
#pragma once
#include "common.cuh"

namespace gnnagg {

struct EdgeParams {
    const int *__restrict__ ptr;
    const int *__restrict__ idx;
    const int *__restrict__ item_row;
    int num_rows;
    int num_edges;
    int num_items;  // ceil(num_edges / kFineItem)
};

enum { kEdgeUAddV = 0, kEdgeWeight = 1, kEdgeDiv = 2 };

// OP = kEdgeUAddV : out[e] = att[2v] + att[2u+1]                      (u_add_v, aggr_gat.h:45)
// OP = kEdgeWeight: out[e] = exp(max(s, slope*s)), s as above          (attGat pass 1, aggr_gat.h:16-18)
// OP = kEdgeDiv   : out[e] = out[e] / center[v]                        (each_div :89 / attGat pass 2 :28)
template <int OP>
__global__ void __launch_bounds__(256) edge_map_kernel(const EdgeParams g, const float *__restrict__ a,
                                                       float *__restrict__ out, float slope)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.num_edges) return;
    const int v = row_of_edge(g.ptr, g.item_row, g.num_items, g.num_rows, e);
    if (OP == kEdgeDiv) {
        out[e] = out[e] / __ldg(a + v);
    } else {
        const float s = __ldg(a + 2 * (size_t)v) + __ldg(a + 2 * (size_t)__ldg(g.idx + e) + 1);
        out[e] = (OP == kEdgeUAddV) ? s : __expf(fmaxf(s, s * slope));
    }
}

// out[v] = sum over the edges e of row v of x_e (add_to_center, aggr_gat.h:50-74; the row-sum half of attGat,
// :19-25).  One thread walks one kRowsumItem-edge item; rows crossing item boundaries leave a partial in
// carry[item] that rowsum_fixup_kernel adds in item order.  Row v is written to out[v * ostride]
// (ostride = 2 fills one column of an [n,2] attention-gradient table).  What x_e is:
//   kSumArray   : in[e]
//   kSumWeights : exp(lrelu(att[2v] + att[2 idx[e] + 1])), the GAT edge weight computed on the fly (no m-float temporary)
enum { kSumArray = 0, kSumWeights = 1 };

// edges per thread of the row-sum kernels: a quarter of the item_row granularity, so that a 1 M-edge graph still
// gives every SM a few hundred threads (one thread per 128 edges left a 148-SM part at 36 CTAs)
constexpr int kRowsumItem = 32;
static_assert(kFineItem % kRowsumItem == 0, "row-sum items must nest inside item_row blocks");

__device__ __forceinline__ int rowsum_start_row(const EdgeParams &g, int e0)
{
    if (e0 % kFineItem == 0) return (e0 == 0) ? 0 : __ldg(g.item_row + e0 / kFineItem);
    return row_of_edge(g.ptr, g.item_row, g.num_items, g.num_rows, e0);
}

template <int SRC>
__global__ void __launch_bounds__(256) rowsum_kernel(const EdgeParams g, const float *__restrict__ in, float slope,
                                                     float *__restrict__ out, float *__restrict__ carry, int ostride)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int64_t)item * kRowsumItem >= g.num_edges) return;
    const int e0 = item * kRowsumItem;
    const int e1 = (int)min((int64_t)g.num_edges, (int64_t)e0 + kRowsumItem);
    int row = rowsum_start_row(g, e0);
    int row_end = __ldg(g.ptr + row + 1);
    bool carry_in = __ldg(g.ptr + row) < e0;
    float acc = 0.f;
    float a_dst = (SRC == kSumWeights) ? __ldg(in + 2 * (size_t)row) : 0.f;
    auto flush = [&]() {
        if (carry_in) {
            carry[item] = acc;
            carry_in = false;
        } else {
            out[(size_t)row * ostride] = acc;
        }
        acc = 0.f;
        ++row;
        row_end = (row < g.num_rows) ? __ldg(g.ptr + row + 1) : INT_MAX;
        if (SRC == kSumWeights && row < g.num_rows) a_dst = __ldg(in + 2 * (size_t)row);
    };
    // the array streamed in edge order: values or source ids; 16-byte aligned at e0 when its base is
    const void *stream = (SRC == kSumArray) ? (const void *)in : (const void *)g.idx;
    const bool vec = ((reinterpret_cast<uintptr_t>(stream) & 15) == 0);
    int e = e0;
    while (row_end == e) flush();
    while (e < e1) {
        uint32_t raw[4];
        const int nb = min(4, e1 - e);
        if (vec && nb == 4) {
            const uint4 t = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(stream) + e));
            raw[0] = t.x, raw[1] = t.y, raw[2] = t.z, raw[3] = t.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) raw[u] = (u < nb) ? __ldg(reinterpret_cast<const uint32_t *>(stream) + e + u) : 0u;
        }
        float src_term[4];  // gathers issued together, ahead of the row walk
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (SRC == kSumWeights)
                src_term[u] = (u < nb) ? __ldg(in + 2 * (size_t)raw[u] + 1) : 0.f;
            else
                src_term[u] = __uint_as_float(raw[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (u < nb) {
                while (row_end == e + u) flush();
                if (SRC == kSumWeights) {
                    const float sc = a_dst + src_term[u];
                    acc += __expf(fmaxf(sc, sc * slope));  // aggr_gat.h:16-18
                } else {
                    acc += src_term[u];
                }
            }
        }
        e += nb;
    }
    if (row_end == e1) {
        while (row < g.num_rows && row_end == e1) flush();
    } else if (carry_in) {
        carry[item] = acc;
    } else {
        out[(size_t)row * ostride] = acc;
    }
}

// two-pass, fixed-order combination of the per-item partials (see agg_fixup_kernel)
constexpr int kRowsumChunk = 256;

template <int PHASE>
__global__ void __launch_bounds__(256) rowsum_fixup_kernel(const EdgeParams g, float *__restrict__ out,
                                                           float *__restrict__ carry, int ostride)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < 1 || (int64_t)item * kRowsumItem >= g.num_edges) return;
    const int e0 = item * kRowsumItem;
    const int row = rowsum_start_row(g, e0);
    const int rs = __ldg(g.ptr + row);
    if (rs >= e0) return;
    const int first = rs / kRowsumItem + 1;
    const int last = (__ldg(g.ptr + row + 1) - 1) / kRowsumItem;
    const bool long_span = (last - first) >= kRowsumChunk;
    int b0, b1, step;
    if (PHASE == 1) {
        if ((item - first) % kRowsumChunk != 0) return;
        b0 = item, b1 = min(last, item + kRowsumChunk - 1), step = 1;
    } else {
        if (item != first || !long_span) return;
        b0 = first, b1 = last, step = kRowsumChunk;
    }
    float acc = 0.f;
    for (int b = b0; b <= b1; b += step) acc += carry[b];
    if (PHASE == 2 || !long_span)
        out[(size_t)row * ostride] += acc;
    else
        carry[item] = acc;
}

// Per-row term of the GAT backward that the edge pass needs, out[v] = (0, c_v = <Y[v,:], dY[v,:]>): 8 lanes per row,
// float4 per lane (the `res` of aggr_gat_fine_bwd, aggr_gat.h:275-283).  The x half is filled by the row-final kernel.
__global__ void __launch_bounds__(256) gat_bwd_rowinfo_kernel(const float *__restrict__ Y, const float *__restrict__ dY,
                                                              float2 *__restrict__ out, int num_rows, int F)
{
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    float acc = 0.f;
    if (row < num_rows)
        for (int col = sub * 4; col < F; col += 32) acc += dot4(ldg_f4(Y + (size_t)row * F + col), ldg_f4(dY + (size_t)row * F + col));
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (row < num_rows && sub == 0) out[row] = make_float2(0.f, acc);
}

// Closes the rows after pass 1 of the GAT backward (agg_kernel<kModeGATBWD>): adds, in a fixed order, the (sum w, sum t)
// partials a row left in the item where it starts (part[v]) and in every EB-edge item it enters (carry[]), then
// rowinfo[v].x = 1 / D_v (the division by `thediv`, aggr_gat.h:244; `den` when the caller hands D_v in) and the
// destination half of the attention gradient d att[2v] = (sum t) / D_v.  8 lanes per row; rows without edges get 0.
__global__ void __launch_bounds__(256) gat_bwd_rowfinal_kernel(const int *__restrict__ ptr, const float2 *__restrict__ part,
                                                               const float2 *__restrict__ carry, const float *__restrict__ den,
                                                               float2 *__restrict__ rowinfo, float *__restrict__ datt,
                                                               int num_rows, int EB)
{
    griddep_wait();
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    float sw = 0.f, st = 0.f;
    if (row < num_rows) {
        const int rs = __ldg(ptr + row), re = __ldg(ptr + row + 1);
        if (re > rs) {
            const int last = (re - 1) / EB;
            for (int b = rs / EB + 1 + sub; b <= last; b += 8) {
                const float2 c = __ldg(carry + b);
                sw += c.x;
                st += c.y;
            }
            if (sub == 0) {
                const float2 c = __ldg(part + row);
                sw += c.x;
                st += c.y;
            }
        }
    }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, off);
        st += __shfl_xor_sync(0xffffffffu, st, off);
    }
    if (row < num_rows && sub == 0) {
        const float d = den ? __ldg(den + row) : sw;
        const float inv = d != 0.f ? __fdividef(1.f, d) : 0.f;
        rowinfo[row].x = inv;
        datt[2 * (size_t)row] = st * inv;
    }
}

// Large graphs, between the two passes of the GAT backward: (w, t) of every edge into transposed edge order and
// normalised on the way, alpha[j] = w_e / D_v and ds[j] = t_e / D_v with e = perm[j], v = t_idx[j].  As a stream of
// independent 8-byte gathers this runs at DRAM speed; fetched inside pass 2 (the small-graph variant, where (w, t) is
// L2 resident) each batch of FMAs would wait for a DRAM round trip.
__global__ void __launch_bounds__(256) gat_bwd_permute_kernel(const float2 *__restrict__ wt, const int *__restrict__ perm,
                                                              const int *__restrict__ t_idx, const float2 *__restrict__ rowinfo,
                                                              float *__restrict__ alpha, float *__restrict__ ds, int count)
{
    griddep_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const float2 v = __ldg(wt + __ldg(perm + j));
    const float inv = __ldg(&rowinfo[__ldg(t_idx + j)].x);
    alpha[j] = v.x * inv;
    ds[j] = v.y * inv;
}

// edge-wise GCN aggregation: Y[dst] += X[src]*val[e], one virtual warp per edge and 128-bit
// reductions (aggr_gcn_edgewise, aggr_gcn.h:291-302: warp per edge, F=32 only, scalar atomics)
template <int LPR>
__global__ void __launch_bounds__(256) gcn_edgewise_kernel(const EdgeParams g, const float *__restrict__ val,
                                                           const float *__restrict__ X, float *__restrict__ Y, int F)
{
    constexpr int VPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int64_t e64 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * VPW + lane / LPR;
    if (e64 >= g.num_edges) return;
    const int e = (int)e64;
    const int dst = row_of_edge(g.ptr, g.item_row, g.num_items, g.num_rows, e);
    const int src = __ldg(g.idx + e);
    const float w = __ldg(val + e);
    for (int col = (lane % LPR) * 4; col < F; col += LPR * 4) {
        const float4 x = ldg_f4(X + (size_t)src * F + col);
        red_add_f4(Y + (size_t)dst * F + col, make_float4(x.x * w, x.y * w, x.z * w, x.w * w));
    }
}

// (src, dst) pairs, edgelist[2e] = idx[e], edgelist[2e+1] = row(e)   (aggregator.h:11-23)
__global__ void __launch_bounds__(256) csr2edgelist_kernel(const EdgeParams g, int *__restrict__ edgelist)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.num_edges) return;
    const int v = row_of_edge(g.ptr, g.item_row, g.num_items, g.num_rows, e);
    reinterpret_cast<int2 *>(edgelist)[e] = make_int2(__ldg(g.idx + e), v);
}

// out[i,:] = X[rows[i],:]: one float4 per thread, F/4 consecutive threads per row (halo packing)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ X, const int64_t *__restrict__ rows,
                                                          float *__restrict__ out, int64_t total4, int F4)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total4) return;
    const int64_t i = t / F4;
    const int c = (int)(t % F4);
    reinterpret_cast<float4 *>(out)[t] = __ldg(reinterpret_cast<const float4 *>(X) + __ldg(rows + i) * F4 + c);
}

// thread-per-row SpMM of include/spmm.h:223-265 (kept for API completeness; rows without edges
// are left untouched exactly as there, :236-237)
__global__ void __launch_bounds__(128) spmm_naive_kernel(int num_v, const int *__restrict__ ptr,
                                                         const int *__restrict__ idx, const float *__restrict__ val,
                                                         const float *__restrict__ X, float *__restrict__ Y, int F)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_v) return;
    const int begin = ptr[r], end = ptr[r + 1];
    if (begin == end) return;
    for (int c = 0; c < F; c += 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = begin; e < end; ++e) fma4(acc, __ldg(val + e), ldg_f4(X + (size_t)__ldg(idx + e) * F + c));
        stg_f4(Y + (size_t)r * F + c, acc);
    }
}

// mismatch counters (spmm.h:11-33)
__global__ void __launch_bounds__(128) validate_kernel(const float *__restrict__ ref, const float *__restrict__ ans,
                                                       int64_t num, int *diffnum)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < num && fabsf((ref[t] - ans[t]) / ref[t]) > 1e-2f) atomicAdd(diffnum, 1);
}

__global__ void __launch_bounds__(128) validate_reordered_kernel(const float *__restrict__ ref,
                                                                 const float *__restrict__ ans,
                                                                 const int *__restrict__ map, int num_v, int F,
                                                                 int *diffnum)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int64_t)num_v * F && fabsf(ref[t] - ans[(size_t)map[t / F] * F + t % F]) > 1e-2f) atomicAdd(diffnum, 1);
}

}  // namespace gnnagg
