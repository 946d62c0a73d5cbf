// dense_tc.cu -- the dense combination C[M,N] = A[M,K] * B[K,N] (row-major fp32) on the 5th-gen
// tensor cores: tcgen05.mma kind::tf32 with the accumulator in TMEM, fp32-level accuracy through
// the 3xTF32 split  A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo.
//
// Replaces matmul_NN (include/dense.h:4-23: cublasSgemm + cublasSgeam transpose) and the
// shuffle-based 32-wide matvec inside aggr_gcn_nn (include/aggr_gcn.h:341-357, OUT <= 32 only).
//
// Shape of the problem: M = #vertices (10^5..10^8), K = feat_in, N = feat_out (32..256).  The GEMM
// is tiny next to the gather-bound aggregation (< 1 % of the layer at reddit shape), so the design
// goal is simplicity and exactness, not tensor peak:
//   * persistent CTA, one 128-row tile of A at a time, B (=W) split once per CTA and resident in
//     shared memory as Bhi/Blo in the K-major no-swizzle canonical UMMA layout;
//   * A streamed in 32-wide K chunks through a 2-stage ring: all 256 threads load fp32, split
//     hi/lo in registers (cvt.rna.tf32) and store the two operands into the canonical layout
//     (TMA cannot be used for the operands because of the split);
//   * one thread issues 12 tcgen05.mma per chunk (4 k-steps x 3 products), tcgen05.commit frees
//     the stage; the epilogue reads TMEM with tcgen05.ld (32 lanes x 32 bit x 16 columns) and
//     stores rows straight to global memory.
#include <cuda_runtime.h>

#include <mutex>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "gnnagg.h"
#include "internal.h"

namespace gnnagg {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                              // floats of K per pipeline stage
constexpr int kStageBytes = kTileM * kChunkK * 4;        // one operand (hi or lo) of one stage: 16 KB
constexpr int kDenseThreads = 256;
constexpr int kMaxNK = 16384;                            // N_slab * K limit: Bhi+Blo <= 128 KB

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) start>>4, [16,30) leading byte offset>>4 (between the two 16-byte K chunks of one MMA),
// [32,46) stride byte offset>>4 (between 8-row core matrices), [46,48) version = 1, swizzle = 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ float tf32_round(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// C[:, n_off : n_off+Ns] = A * B[:, n_off : n_off+Ns];  lda = K, ldb = ldc = N (row-major)
__global__ void __launch_bounds__(kDenseThreads, 1)
dense_tf32x3_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int64_t M, int N,
                    int K, int n_off, int Ns, int tmem_cols)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_empty[2];
    __shared__ __align__(8) uint64_t s_acc;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_bytes = (uint32_t)Ns * K * 4;  // one of Bhi / Blo
    uint8_t *sBhi = smem, *sBlo = smem + b_bytes, *sA = smem + 2 * b_bytes;  // sA: [stage][hi|lo][16 KB]
    const uint32_t sbo_b = (uint32_t)K * 32;                                   // 8 rows of B^T: (K/4) chunks * 128 B
    const uint32_t sbo_a = (kChunkK / 4) * 128;                                // 1024 B

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(smem_u32(&s_empty[0]), 1);
        mbar_init(smem_u32(&s_empty[1]), 1);
        mbar_init(smem_u32(&s_acc), 1);
        fence_mbar_init();
    }
    // B^T (n-major rows, K contiguous) split into hi / lo, canonical K-major layout:
    //   off(n,k) = (n/8)*sbo_b + (k/4)*128 + (n%8)*16 + (k%4)*4
    for (int i = tid; i < Ns * K; i += kDenseThreads) {
        const int k = i / Ns, n = i % Ns;  // coalesced along n
        const float w = __ldg(B + (size_t)k * N + n_off + n);
        const float hi = tf32_round(w);
        const uint32_t off = (uint32_t)(n >> 3) * sbo_b + (uint32_t)(k >> 2) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 3) * 4u;
        *reinterpret_cast<float *>(sBhi + off) = hi;
        *reinterpret_cast<float *>(sBlo + off) = w - hi;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    // instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Ns >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const int chunks = K / kChunkK;
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    uint32_t c = 0;          // chunks produced so far by this CTA (ring position)
    uint32_t acc_phase = 0;  // parity of s_acc

    // this thread's 4 float4 of chunk (tile, kc): 128 rows x 8 float4; a warp covers 8 rows x 64 B
    // (full sectors from global, conflict-free 128-byte runs in shared memory)
    auto fetch = [&](int64_t tile, int kc, float4 (&v)[4]) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int u = p * kDenseThreads + tid;
            const int r8 = u & 7, kq = (u >> 3) & 7, rb = u >> 6;
            const int64_t row = tile * kTileM + rb * 8 + r8;
            v[p] = (tile < tiles && row < M) ? ldg_f4(A + (size_t)row * K + kc * kChunkK + kq * 4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 cur[4], nxt[4];
    fetch(blockIdx.x, 0, cur);

    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        for (int kc = 0; kc < chunks; ++kc, ++c) {
            // software pipeline: the global loads of the NEXT chunk are in flight while this one is
            // split, stored and multiplied
            if (kc + 1 < chunks)
                fetch(tile, kc + 1, nxt);
            else
                fetch(tile + gridDim.x, 0, nxt);
            const uint32_t stage = c & 1;
            if (c >= 2) mbar_wait(smem_u32(&s_empty[stage]), ((c >> 1) - 1) & 1);  // MMAs of chunk c-2 retired
            uint8_t *aHi = sA + stage * 2 * kStageBytes, *aLo = aHi + kStageBytes;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int u = p * kDenseThreads + tid;
                const int r8 = u & 7, kq = (u >> 3) & 7, rb = u >> 6;
                const float4 v = cur[p];
                const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                const uint32_t off = (uint32_t)rb * sbo_a + (uint32_t)kq * 128u + (uint32_t)r8 * 16u;
                *reinterpret_cast<float4 *>(aHi + off) = h;
                *reinterpret_cast<float4 *>(aLo + off) = l;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) cur[p] = nxt[p];
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_hi = smem_u32(aHi), a_lo = smem_u32(aLo);
                const uint32_t b_hi = smem_u32(sBhi) + (uint32_t)kc * (kChunkK / 4) * 128u;
                const uint32_t b_lo = smem_u32(sBlo) + (uint32_t)kc * (kChunkK / 4) * 128u;
#pragma unroll
                for (int j = 0; j < kChunkK / 8; ++j) {  // one MMA consumes K = 8 tf32 = two 16-byte chunks
                    const uint64_t dah = smem_desc(a_hi + j * 256, 128, sbo_a), dal = smem_desc(a_lo + j * 256, 128, sbo_a);
                    const uint64_t dbh = smem_desc(b_hi + j * 256, 128, sbo_b), dbl = smem_desc(b_lo + j * 256, 128, sbo_b);
                    tc_mma_tf32(tmem, dal, dbh, idesc, (kc | j) ? 1u : 0u);  // small terms first
                    tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
                    tc_mma_tf32(tmem, dah, dbh, idesc, 1u);
                }
                tc_commit(smem_u32(&s_empty[stage]));
                if (kc == chunks - 1) tc_commit(smem_u32(&s_acc));
            }
        }
        // ---- epilogue: TMEM -> registers -> global.  warp w reads lanes 32*(w%4).., column half w/4
        mbar_wait(smem_u32(&s_acc), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        {
            const int q = warp & 3, half = warp >> 2;
            const int64_t row = row0 + q * 32 + lane;
            const int cbeg = half * (Ns / 2), cend = cbeg + Ns / 2;
            for (int col = cbeg; col < cend; col += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
                    float *dst = C + (size_t)row * N + n_off + col;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        stg_f4(dst + i, make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                    __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM may be overwritten by the next tile's first MMA
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Streamed-W variant for feat_in*feat_out > 16 384 (W's hi/lo split no longer fits beside the A ring).
// W is split ONCE per call into Whi / Wlo laid out chunk by chunk in the canonical operand layout
// (split_w_kernel); the main kernel then streams BOTH operands through a 2-stage ring in 32-wide K chunks: the A
// chunk is produced by the threads (fp32 -> hi/lo, as above), the two W chunks arrive by bulk copy (TMA, UBLKCP)
// from L2 on the stage's `full` mbarrier while the threads convert A.  One launch covers all N columns, so A is
// read once (the resident variant would need N/64 column slabs and re-read A for each).
// ---------------------------------------------------------------------------------------------------------
// element (k, n) of W -> chunk kc = k/32:  kc*N*32 + (n/8)*256 + ((k%32)/4)*32 + (n%8)*4 + (k%4)   [floats]
__global__ void __launch_bounds__(256) split_w_kernel(const float *__restrict__ W, float *__restrict__ Whi,
                                                      float *__restrict__ Wlo, int K, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const int k = i / N, n = i % N;
    const float w = __ldg(W + i);
    const float hi = tf32_round(w);
    const size_t off = (size_t)(k >> 5) * N * 32 + (size_t)(n >> 3) * 256 + (size_t)((k & 31) >> 2) * 32 + (size_t)(n & 7) * 4 + (k & 3);
    Whi[off] = hi;
    Wlo[off] = w - hi;
}

__global__ void __launch_bounds__(kDenseThreads, 1)
dense_tf32x3_stream_kernel(const float *__restrict__ A, const float *__restrict__ Whi, const float *__restrict__ Wlo,
                           float *__restrict__ C, int64_t M, int N, int K, int tmem_cols)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_empty[2], s_full[2];
    __shared__ __align__(8) uint64_t s_acc;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_chunk = (uint32_t)N * 128u;                 // one of Whi / Wlo, one K chunk
    const uint32_t stage_bytes = 2u * kStageBytes + 2u * b_chunk;  // [Ahi | Alo | Bhi | Blo]
    const uint32_t sbo = (kChunkK / 4) * 128;                    // 1024 B between 8-row core-matrix groups, both operands

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s_empty[i]), 1);
            mbar_init(smem_u32(&s_full[i]), 1);
        }
        mbar_init(smem_u32(&s_acc), 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const int chunks = K / kChunkK;
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    uint32_t c = 0, acc_phase = 0;

    auto fetch = [&](int64_t tile, int kc, float4 (&v)[4]) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int u = p * kDenseThreads + tid;
            const int r8 = u & 7, kq = (u >> 3) & 7, rb = u >> 6;
            const int64_t row = tile * kTileM + rb * 8 + r8;
            v[p] = (tile < tiles && row < M) ? ldg_f4(A + (size_t)row * K + kc * kChunkK + kq * 4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 cur[4], nxt[4];
    fetch(blockIdx.x, 0, cur);

    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kTileM;
        for (int kc = 0; kc < chunks; ++kc, ++c) {
            if (kc + 1 < chunks)
                fetch(tile, kc + 1, nxt);
            else
                fetch(tile + gridDim.x, 0, nxt);
            const uint32_t stage = c & 1;
            if (c >= 2) mbar_wait(smem_u32(&s_empty[stage]), ((c >> 1) - 1) & 1);  // MMAs of chunk c-2 retired
            uint8_t *aHi = smem + stage * stage_bytes, *aLo = aHi + kStageBytes;
            uint8_t *bHi = aLo + kStageBytes, *bLo = bHi + b_chunk;
            if (tid == 0) {  // the W chunks of this K step travel while the threads convert A
                const uint32_t full = smem_u32(&s_full[stage]);
                mbar_expect_tx(full, 2u * b_chunk);
                bulk_g2s(smem_u32(bHi), Whi + (size_t)kc * N * 32, b_chunk, full);
                bulk_g2s(smem_u32(bLo), Wlo + (size_t)kc * N * 32, b_chunk, full);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int u = p * kDenseThreads + tid;
                const int r8 = u & 7, kq = (u >> 3) & 7, rb = u >> 6;
                const float4 v = cur[p];
                const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
                const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                const uint32_t off = (uint32_t)rb * sbo + (uint32_t)kq * 128u + (uint32_t)r8 * 16u;
                *reinterpret_cast<float4 *>(aHi + off) = h;
                *reinterpret_cast<float4 *>(aLo + off) = l;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) cur[p] = nxt[p];
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                mbar_wait(smem_u32(&s_full[stage]), (c >> 1) & 1);  // W chunks landed
                tc_fence_after();
                const uint32_t a_hi = smem_u32(aHi), a_lo = smem_u32(aLo), b_hi = smem_u32(bHi), b_lo = smem_u32(bLo);
#pragma unroll
                for (int j = 0; j < kChunkK / 8; ++j) {
                    const uint64_t dah = smem_desc(a_hi + j * 256, 128, sbo), dal = smem_desc(a_lo + j * 256, 128, sbo);
                    const uint64_t dbh = smem_desc(b_hi + j * 256, 128, sbo), dbl = smem_desc(b_lo + j * 256, 128, sbo);
                    tc_mma_tf32(tmem, dal, dbh, idesc, (kc | j) ? 1u : 0u);
                    tc_mma_tf32(tmem, dah, dbl, idesc, 1u);
                    tc_mma_tf32(tmem, dah, dbh, idesc, 1u);
                }
                tc_commit(smem_u32(&s_empty[stage]));
                if (kc == chunks - 1) tc_commit(smem_u32(&s_acc));
            }
        }
        mbar_wait(smem_u32(&s_acc), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        {
            const int q = warp & 3, half = warp >> 2;
            const int64_t row = row0 + q * 32 + lane;
            const int cbeg = half * (N / 2), cend = cbeg + N / 2;
            for (int col = cbeg; col < cend; col += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
                    float *dst = C + (size_t)row * N + col;
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        stg_f4(dst + i, make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                    __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
}

// Stream-ordered pool for the W split of the streamed kernel.  The device's default pool hands freed memory back to
// the driver at every synchronisation (release threshold 0), which made each call re-map its 2*K*N floats; a private
// pool that keeps what it has turns the allocation into a pointer bump.  One pool per device, created on first use.
static cudaMemPool_t split_pool(int dev)
{
    static std::mutex lock;
    static cudaMemPool_t pools[64] = {};
    if (dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> guard(lock);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) {
            pools[dev] = nullptr;
            return nullptr;
        }
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return pools[dev];
}

int dense_nn_launch(const float *A, const float *B, float *C, int64_t M, int N, int K, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (N < 32 || K < 32 || N > 256 || K > 256 || (N % 32) || (K % 32))
        return set_error(GNNAGG_ERR_ARG, "dense combination: feat_in and feat_out must be multiples of 32 in [32,256]");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C)) & 15)
        return set_error(GNNAGG_ERR_ARG, "dense combination: A and C must be 16-byte aligned");
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return set_error(GNNAGG_ERR_CUDA, "dense combination: no CUDA device");
    const int64_t tiles_all = (M + kTileM - 1) / kTileM;
    if (N * K > kMaxNK) {
        // W's split does not fit in shared memory next to the A ring: stream it (see dense_tf32x3_stream_kernel)
        float *wsplit = nullptr;
        cudaMemPool_t pool = split_pool(dev);
        if (!pool || cudaMallocFromPoolAsync((void **)&wsplit, (size_t)2 * K * N * sizeof(float), pool, st) != cudaSuccess)
            return set_error(GNNAGG_ERR_CUDA, "dense combination: cannot allocate the W split");
        float *whi = wsplit, *wlo = wsplit + (size_t)K * N;
        split_w_kernel<<<(K * N + 255) / 256, 256, 0, st>>>(B, whi, wlo, K, N);
        int cols = 32;
        while (cols < N) cols *= 2;
        const size_t smem_s = 2 * ((size_t)2 * kStageBytes + (size_t)2 * N * 128);
        static size_t configured_s = 0;
        if (smem_s > configured_s) {
            if (cudaFuncSetAttribute(dense_tf32x3_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s) !=
                cudaSuccess)
                return set_error(GNNAGG_ERR_CUDA, "dense combination: cannot raise dynamic shared memory");
            configured_s = smem_s;
        }
        const unsigned grid_s = (unsigned)(tiles_all < sms ? tiles_all : sms);
        dense_tf32x3_stream_kernel<<<grid_s, kDenseThreads, smem_s, st>>>(A, whi, wlo, C, M, N, K, cols);
        const cudaError_t e = cudaPeekAtLastError();
        cudaFreeAsync(wsplit, st);
        return e == cudaSuccess ? GNNAGG_OK : set_error(GNNAGG_ERR_CUDA, cudaGetErrorString(e));
    }
    const int Ns = N;  // W resident in shared memory
    int tmem_cols = 32;
    while (tmem_cols < Ns) tmem_cols *= 2;
    const size_t smem = (size_t)2 * Ns * K * 4 + 4 * kStageBytes;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(dense_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_error(GNNAGG_ERR_CUDA, "dense combination: cannot raise dynamic shared memory");
        configured = smem;
    }
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    dense_tf32x3_kernel<<<grid, kDenseThreads, smem, st>>>(A, B, C, M, N, K, 0, Ns, tmem_cols);
    return cudaPeekAtLastError() == cudaSuccess ? GNNAGG_OK : set_error(GNNAGG_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
}

}  // namespace gnnagg
