// sample_device.cu -- the sub-graph samplers of include/sample.h (sampleVertex :131-200,
// sampleVertexSampleNeighbor :274-357) built for the GPU: the step in front of the aggregation in a
// mini-batch pipeline; their CSRSubGraph output feeds the same aggregators.
//
// sampleVertex is deterministic and reproduced bit for bit: the active set grows by whole neighbourhoods
// for layer_num - 1 hops (expandActive, :109-124), then the rows of the active vertices are extracted in
// ascending vertex order with their complete neighbour lists and GLOBAL source ids (:150-199).
// Differences in how: the hop is edge-parallel (the reference gives a hub row to one warp), the row copy is
// edge-parallel (moveEdge, :59-75, is warp-per-row), scans/compaction come from CUB instead of four thrust calls
// and two 1-thread kernels, and there is one host synchronisation (two counters) instead of four.
//
// sampleVertexSampleNeighbor is RE-SPECIFIED: the reference draws curand() % deg until `limit` unmarked
// positions are hit, marks them in an m-int `chosen` array, and then copies rows with moveEdgeSelective whose
// mark test is inverted for deg >= 2*limit and matches nothing for limit < deg < 2*limit (:97-104 vs :229-246) --
// it writes past the row or leaves it uninitialised, and no driver calls it.  Here a row longer than the fanout
// contributes exactly `fanout` distinct neighbours by stratified sampling: position j of the sample is drawn
// uniformly from the j-th of `fanout` equal strata of the row (so CSR order is kept and every edge has inclusion
// probability fanout/deg), as a pure function of (seed, vertex, j) -- no per-vertex generator state, no m-sized
// marker array, identical on the CPU oracle (oracle.c: orc_sample_pos) and on any GPU, and a vertex reached in
// several hops keeps one sample, which is what the reference's `expanded` flags are for (:217-219).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"
#include "gnnagg.h"
#include "internal.h"

namespace gnnagg {

#define SP_TRY(expr)                                            \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) {                                \
            set_error(GNNAGG_ERR_CUDA, cudaGetErrorString(_e)); \
            goto fail;                                          \
        }                                                       \
    } while (0)

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// position (offset inside the row) of the j-th sampled neighbour of vertex v, deg > fanout
__host__ __device__ __forceinline__ int sample_pos(uint64_t seed, int v, int j, int deg, int fanout)
{
    const int lo = (int)(((int64_t)j * deg) / fanout);
    const int hi = (int)(((int64_t)(j + 1) * deg) / fanout);
    const uint64_t r = splitmix64(seed ^ ((uint64_t)(uint32_t)v * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)j * 0xD1B54A32D192ED03ull));
    return lo + (int)(r % (uint64_t)(hi - lo));
}

__global__ void __launch_bounds__(256) normalise_flags_kernel(int *__restrict__ active, int n)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) active[v] = active[v] != 0;
}

// one hop over complete neighbourhoods: out[u] = 1 for every edge (v <- u) with v active (out starts as a copy of active)
__global__ void __launch_bounds__(256) expand_full_kernel(const int *__restrict__ ptr, const int *__restrict__ idx,
                                                          const int *__restrict__ item_row, int num_items, int n, int m,
                                                          const int *__restrict__ active, int *__restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    if (__ldg(active + row_of_edge(ptr, item_row, num_items, n, e))) out[__ldg(idx + e)] = 1;
}

// one hop over sampled neighbourhoods
__global__ void __launch_bounds__(256) expand_sampled_kernel(const int *__restrict__ ptr, const int *__restrict__ idx, int n,
                                                             const int *__restrict__ active, int *__restrict__ out, int fanout,
                                                             uint64_t seed)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || !__ldg(active + v)) return;
    const int begin = __ldg(ptr + v), deg = __ldg(ptr + v + 1) - begin;
    if (deg <= fanout) {
        for (int i = 0; i < deg; ++i) out[__ldg(idx + begin + i)] = 1;
    } else {
        for (int j = 0; j < fanout; ++j) out[__ldg(idx + begin + sample_pos(seed, v, j, deg, fanout))] = 1;
    }
}

__global__ void __launch_bounds__(256) sub_degree_kernel(const int *__restrict__ ptr, const int *__restrict__ vertexset, int count,
                                                         int fanout, int *__restrict__ deg)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > count) return;
    int d = 0;
    if (r < count) {
        const int v = __ldg(vertexset + r);
        d = __ldg(ptr + v + 1) - __ldg(ptr + v);
        if (fanout > 0 && d > fanout) d = fanout;
    }
    deg[r] = d;  // deg[count] = 0 so that the exclusive scan over count+1 entries ends with the edge total
}

// one thread per output edge
__global__ void __launch_bounds__(256) sub_fill_kernel(const int *__restrict__ ptr, const int *__restrict__ idx,
                                                       const int *__restrict__ vertexset, const int *__restrict__ sub_ptr,
                                                       const int *__restrict__ sub_item_row, int sub_items, int count,
                                                       int sub_edges, int fanout, uint64_t seed, int *__restrict__ sub_idx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sub_edges) return;
    const int r = row_of_edge(sub_ptr, sub_item_row, sub_items, count, j);
    const int v = __ldg(vertexset + r);
    const int begin = __ldg(ptr + v);
    int off = j - __ldg(sub_ptr + r);
    if (fanout > 0) {
        const int deg = __ldg(ptr + v + 1) - begin;
        if (deg > fanout) off = sample_pos(seed, v, off, deg, fanout);
    }
    sub_idx[j] = __ldg(idx + begin + off);
}

__global__ void __launch_bounds__(256) sub_item_row_kernel(const int *__restrict__ ptr, int num_rows, int num_edges,
                                                           int *__restrict__ item_row, int num_items)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= num_items) return;
    const int64_t e64 = (int64_t)k * kFineItem;
    const int e = (int)(e64 < num_edges ? e64 : num_edges - 1);
    int lo = 0, hi = num_rows - 1;  // last r with ptr[r] <= e
    while (lo < hi) {
        const int mid = (int)(((int64_t)lo + hi + 1) >> 1);
        if (__ldg(ptr + mid) <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    item_row[k] = lo;
}

static inline unsigned blocks(int64_t n) { return (unsigned)((n + 255) / 256); }

// fanout <= 0: complete neighbourhoods (sampleVertex); fanout > 0: at most `fanout` neighbours per row.
// d_active [n] is updated in place to the expanded 0/1 set (the reference does the same through `int *&`).
// Outputs are cudaMalloc'ed and owned by the caller.
int sample_subgraph_device(const int *d_ptr, const int *d_idx, const int *d_item_row, int num_items, int n, int m,
                           int *d_active, int fanout, int layer_num, uint64_t seed, int **vertexset, int **sub_ptr,
                           int **sub_idx, int *num_v, int *num_e, cudaStream_t st)
{
    *vertexset = *sub_ptr = *sub_idx = nullptr;
    *num_v = *num_e = 0;
    int *other = nullptr, *count_d = nullptr, *deg = nullptr, *sub_item_row = nullptr;
    void *tmp = nullptr;
    size_t need = 0, tmp_bytes = 0;
    int count = 0, edges = 0, sub_items = 0;
    thrust::counting_iterator<int> ids(0);
    auto scratch = [&](size_t bytes) -> cudaError_t {
        if (bytes <= tmp_bytes) return cudaSuccess;
        if (tmp) cudaFree(tmp);
        tmp = nullptr;
        tmp_bytes = 0;
        cudaError_t e = cudaMalloc(&tmp, bytes ? bytes : 1);
        if (e == cudaSuccess) tmp_bytes = bytes;
        return e;
    };
    const size_t nn = (size_t)(n > 0 ? n : 1);

    if (n > 0) normalise_flags_kernel<<<blocks(n), 256, 0, st>>>(d_active, n);
    if (layer_num > 1 && n > 0) {
        SP_TRY(cudaMalloc((void **)&other, nn * sizeof(int)));
        for (int hop = 0; hop < layer_num - 1; ++hop) {
            SP_TRY(cudaMemcpyAsync(other, d_active, nn * sizeof(int), cudaMemcpyDeviceToDevice, st));
            if (fanout > 0)
                expand_sampled_kernel<<<blocks(n), 256, 0, st>>>(d_ptr, d_idx, n, d_active, other, fanout, seed);
            else if (m > 0)
                expand_full_kernel<<<blocks(m), 256, 0, st>>>(d_ptr, d_idx, d_item_row, num_items, n, m, d_active, other);
            SP_TRY(cudaMemcpyAsync(d_active, other, nn * sizeof(int), cudaMemcpyDeviceToDevice, st));
        }
    }
    // compaction of the active ids (ascending), their degrees, the new row pointers
    SP_TRY(cudaMalloc((void **)&count_d, sizeof(int)));
    SP_TRY(cudaMalloc((void **)vertexset, nn * sizeof(int)));
    SP_TRY(cub::DeviceSelect::Flagged(nullptr, need, ids, d_active, *vertexset, count_d, n, st));
    SP_TRY(scratch(need));
    SP_TRY(cub::DeviceSelect::Flagged(tmp, tmp_bytes, ids, d_active, *vertexset, count_d, n, st));
    SP_TRY(cudaMemcpyAsync(&count, count_d, sizeof(int), cudaMemcpyDeviceToHost, st));
    SP_TRY(cudaStreamSynchronize(st));
    SP_TRY(cudaMalloc((void **)&deg, ((size_t)count + 1) * sizeof(int)));
    SP_TRY(cudaMalloc((void **)sub_ptr, ((size_t)count + 1) * sizeof(int)));
    sub_degree_kernel<<<blocks((int64_t)count + 1), 256, 0, st>>>(d_ptr, *vertexset, count, fanout, deg);
    SP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, deg, *sub_ptr, count + 1, st));
    SP_TRY(scratch(need));
    SP_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, *sub_ptr, count + 1, st));
    SP_TRY(cudaMemcpyAsync(&edges, *sub_ptr + count, sizeof(int), cudaMemcpyDeviceToHost, st));
    SP_TRY(cudaStreamSynchronize(st));
    SP_TRY(cudaMalloc((void **)sub_idx, (size_t)(edges > 0 ? edges : 1) * sizeof(int)));
    if (edges > 0) {
        sub_items = (int)(((int64_t)edges + kFineItem - 1) / kFineItem);
        SP_TRY(cudaMalloc((void **)&sub_item_row, (size_t)sub_items * sizeof(int)));
        sub_item_row_kernel<<<blocks(sub_items), 256, 0, st>>>(*sub_ptr, count, edges, sub_item_row, sub_items);
        sub_fill_kernel<<<blocks(edges), 256, 0, st>>>(d_ptr, d_idx, *vertexset, *sub_ptr, sub_item_row, sub_items, count, edges,
                                                      fanout, seed, *sub_idx);
    }
    SP_TRY(cudaGetLastError());
    SP_TRY(cudaStreamSynchronize(st));
    cudaFree(other), cudaFree(count_d), cudaFree(deg), cudaFree(sub_item_row), cudaFree(tmp);
    *num_v = count;
    *num_e = edges;
    return GNNAGG_OK;
fail:
    cudaFree(other), cudaFree(count_d), cudaFree(deg), cudaFree(sub_item_row), cudaFree(tmp);
    cudaFree(*vertexset), cudaFree(*sub_ptr), cudaFree(*sub_idx);
    *vertexset = *sub_ptr = *sub_idx = nullptr;
    return GNNAGG_ERR_CUDA;
}

}  // namespace gnnagg
