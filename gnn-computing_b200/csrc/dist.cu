// dist.cu -- multi-GPU aggregation over NVLink peer memory (gnnagg_dist_*).
//
// The reference is single-GPU: every driver asserts GPUNUM == 1 (Figure9/main.cu:19) and the multi-GPU
// leftovers (GPUNUM/gptrs/gidxs globals, syncAll: include/util.h:39-57,135-142; prepare*Multi prototypes:
// include/data.h:48-58) have no implementation.  What is built here: the graph is 1-D partitioned by
// destination row; rank r owns a row block and the matching shard of X.  The only exchange step is the
// source-feature halo.  It is NOT a library collective:
//
//   * every rank keeps its X shard in a peer-visible allocation (same process: peer access; one process per
//     GPU: cudaIpc handles exchanged once at set-up);
//   * at set-up each rank finds the distinct REMOTE source rows its block references (mark / scan / fill on the
//     GPU), re-indexes its CSR into [own shard | compact receive buffer] coordinates and splits it by the OWNER
//     of the source into stages (the slices of locality_schedule, graph_schedule.h:24-29, kept as CSRs);
//   * a step: stage 0 (edges whose source is local) is aggregated straight from the shard while
//     halo_pull_kernel -- plain 128-bit loads from the owners' shards over NVLink, no pack kernel, no send
//     buffer, no NCCL -- fills the receive buffer owner by owner on a high-priority stream, every rank
//     starting with a different owner; stage s is accumulated (gnnagg_gcn_run_acc, deterministic) as soon as
//     its owners have landed.  Cross-GPU ordering is a pair of epoch flags per peer in peer-visible memory:
//     ready[p] ("p's shard holds the X of epoch e": release store after the producer, acquire spin in the
//     puller) and done[p] ("p has finished reading my shard for epoch e", write-after-read protection for
//     the next producer).  Spins are bounded (globaltimer) and report through an error word instead of hanging.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "gnnagg.h"
#include "internal.h"

namespace gnnagg {

constexpr int kMaxWorld = GNNAGG_DIST_MAX_WORLD;
constexpr uint64_t kSpinLimitNs = 20ull * 1000 * 1000 * 1000;  // a peer that does not answer within 20 s is reported

struct HaloFlags {
    uint32_t ready[kMaxWorld];  // ready[p]: epoch of the X that rank p's shard currently holds (written by p)
    uint32_t done[kMaxWorld];   // done[p]: last epoch rank p finished pulling from MY shard (written by p)
    uint32_t err;               // first protocol error seen by a kernel of this rank (0 = none)
    uint32_t pad[31];
};

struct PeerTable {
    HaloFlags *flags[kMaxWorld];
};
struct Bounds {
    int64_t b[kMaxWorld + 1];
    int stage_of[kMaxWorld];
    int world;
};

#define DT_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            char _buf[512];                                                                                   \
            snprintf(_buf, sizeof _buf, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return set_error(GNNAGG_ERR_CUDA, _buf);                                                          \
        }                                                                                                     \
    } while (0)

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// waits until *flag has reached `epoch`; gives up after kSpinLimitNs and records `code` in *err
__device__ __forceinline__ bool spin_until(const uint32_t *flag, uint32_t epoch, uint32_t *err, uint32_t code)
{
    if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    const uint64_t t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(128);
        if (global_ns() - t0 > kSpinLimitNs) {
            atomicCAS(err, 0u, code);
            return false;
        }
    }
    return true;
}

// "my shard holds the X of `epoch`": first make sure every peer has finished reading the previous contents
// (done flags in MY memory), then publish the epoch into every peer's ready[rank]
__global__ void __launch_bounds__(32) halo_signal_kernel(HaloFlags *mine, PeerTable peers, int rank, int world, uint32_t epoch)
{
    const int p = threadIdx.x;
    if (p < world && p != rank) spin_until(&mine->done[p], epoch - 1u, &mine->err, 0x100u + (uint32_t)p);
    __syncwarp();
    __threadfence_system();
    if (p < world) st_release_sys(&peers.flags[p]->ready[rank], epoch);
}

// "I have finished reading everybody's shard for `epoch`"
__global__ void __launch_bounds__(32) halo_done_kernel(PeerTable peers, int rank, int world, uint32_t epoch)
{
    const int p = threadIdx.x;
    __threadfence_system();
    if (p < world && p != rank) st_release_sys(&peers.flags[p]->done[rank], epoch);
}

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// dst[i, :] = src[rows[i], :] for the `count` receive slots of one owner; src is that owner's shard, mapped peer
// memory: the loads travel over NVLink (or stay local when the "peer" lives on the same device).  U independent
// 128-bit loads per thread are issued before the first store: 2 CTAs/SM x 256 threads x U x 16 B keep ~64 KB per
// SM in flight, several times the NVLink latency-bandwidth product.
template <int U>
__global__ void __launch_bounds__(256) halo_pull_kernel(const float4 *__restrict__ src, const int *__restrict__ rows,
                                                        float4 *__restrict__ dst, int64_t count4, int F4, int f4_shift,
                                                        const uint32_t *ready, uint32_t epoch, uint32_t *err, uint32_t code)
{
    if (ready != nullptr) {
        __shared__ int ok;
        if (threadIdx.x == 0) ok = spin_until(ready, epoch, err, code) ? 1 : 0;
        __syncthreads();
        if (!ok) return;  // reported through the error word; the receive buffer keeps its old contents
    }
    const int64_t stride = (int64_t)gridDim.x * 256 * U;
    for (int64_t base = (int64_t)blockIdx.x * 256 * U + threadIdx.x; base < count4; base += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * 256;
            if (e < count4) {
                const int64_t r = (f4_shift >= 0) ? (e >> f4_shift) : (e / F4);
                const int c = (int)(e - r * F4);
                v[u] = ld_stream_f4(src + (int64_t)__ldg(rows + r) * F4 + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * 256;
            if (e < count4) dst[e] = v[u];
        }
    }
}

// ---- set-up kernels: which remote rows does this block reference, and where do they land -------------------
__device__ __forceinline__ int owner_of(const Bounds &b, int64_t g)
{
    int p = 0;
    while (p + 1 < b.world && g >= b.b[p + 1]) ++p;
    return p;
}

__global__ void __launch_bounds__(256) dist_mark_kernel(const int *__restrict__ idx, int64_t m, int64_t own_lo, int64_t own_hi,
                                                        int64_t total, int *__restrict__ mark, int *__restrict__ bad)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int64_t g = __ldg(idx + e);
    if (g < 0 || g >= total)
        *bad = 1;
    else if (g < own_lo || g >= own_hi)
        mark[g] = 1;
}

__global__ void __launch_bounds__(256) dist_fill_recv_kernel(const int *__restrict__ mark, const int *__restrict__ pos, int64_t total,
                                                             Bounds b, int *__restrict__ recv_local)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total || !mark[g]) return;
    recv_local[pos[g]] = (int)(g - b.b[owner_of(b, g)]);
}

// new index of every edge (local row of the own shard, or slot of the receive buffer) and its stage
__global__ void __launch_bounds__(256) dist_reindex_kernel(const int *__restrict__ idx, int64_t m, const int *__restrict__ pos,
                                                           int64_t own_lo, int64_t own_hi, Bounds b, int *__restrict__ idx_new,
                                                           int *__restrict__ keys)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const int64_t g = __ldg(idx + e);
    if (g >= own_lo && g < own_hi) {
        idx_new[e] = (int)(g - own_lo);
        keys[e] = 0;
    } else {
        idx_new[e] = __ldg(pos + g);
        keys[e] = b.stage_of[owner_of(b, g)];
    }
}

__global__ void __launch_bounds__(256) dist_gather_val_kernel(const float *__restrict__ val, const int *__restrict__ perm,
                                                              float *__restrict__ out, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = __ldg(val + __ldg(perm + i));
}

static inline unsigned nblocks(int64_t n) { return (unsigned)((n + 255) / 256); }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace gnnagg

using namespace gnnagg;

constexpr int kMaxStages = kMaxWorld;

// what a rank publishes to its peers (gnnagg_dist_export): fits GNNAGG_DIST_BLOB_BYTES
struct DistBlob {
    uint32_t magic;
    int rank, world, device, pid;
    int feat_cap;
    int64_t rows;
    uint64_t raw_base;  // usable as is inside the exporting process
    uint64_t bytes;
    uint64_t off_x[2];
    cudaIpcMemHandle_t mem;
};
static_assert(sizeof(DistBlob) <= GNNAGG_DIST_BLOB_BYTES, "blob too large");

struct gnnagg_dist {
    int rank = 0, world = 1, device = 0, pid = 0;
    int64_t bounds[kMaxWorld + 1] = {0};
    int rows = 0;  // rows of the own shard = rows of the CSR block
    int feat_cap = 0;
    // peer-visible allocation: [flags | x0 | x1]
    char *base = nullptr;
    size_t bytes = 0, off_x[2] = {0, 0};
    // peers as seen from this rank
    char *peer_base[kMaxWorld] = {nullptr};
    size_t peer_off_x[kMaxWorld][2] = {{0, 0}};
    bool peer_ipc[kMaxWorld] = {false};
    bool connected = false;
    // plan
    int64_t num_e = 0;
    int remote_stages = 0, num_stages = 0;  // stage 0 = local sources; 1..remote_stages = groups of owners
    int stage_of[kMaxWorld] = {0};
    int pull_order[kMaxWorld] = {0};  // the world-1 remote owners in the order they are pulled
    int64_t num_recv = 0;
    int recv_off[kMaxWorld + 1] = {0};
    int *recv_local = nullptr;
    float *recv_buf = nullptr;
    // sub-CSRs, one per stage
    int *sl_ptr = nullptr, *sl_idx = nullptr, *sl_perm = nullptr;
    float *sl_val = nullptr;
    int sl_off[kMaxStages] = {0}, sl_cnt[kMaxStages] = {0};
    gnnagg_aggregator *stage[kMaxStages] = {nullptr};
    float *ax = nullptr;
    size_t ax_cap = 0;
    // step machinery
    cudaStream_t comm = nullptr;
    cudaEvent_t ev_sig = nullptr, ev_done = nullptr, ev_stage[kMaxStages] = {nullptr};
    cudaEvent_t t_m0 = nullptr, t_m1 = nullptr, t_m2 = nullptr, t_m3 = nullptr, t_c0 = nullptr, t_c1 = nullptr;
    bool prof = false;
    uint32_t epoch = 0;
    int sm_count = 148;
    int prepared_feat = 0;
    int same_device_ranks = 1;  // ranks (including this one) living on this rank's device: > 1 only in single-GPU tests
    int64_t launches = 0;
};

static HaloFlags *flags_of(char *base) { return reinterpret_cast<HaloFlags *>(base); }

static void free_graph(gnnagg_dist *d)
{
    for (int s = 0; s < kMaxStages; ++s) {
        if (d->stage[s]) gnnagg_destroy(d->stage[s]);
        d->stage[s] = nullptr;
    }
    cudaFree(d->sl_ptr), cudaFree(d->sl_idx), cudaFree(d->sl_perm), cudaFree(d->sl_val);
    cudaFree(d->recv_local), cudaFree(d->recv_buf);
    d->sl_ptr = d->sl_idx = d->sl_perm = d->recv_local = nullptr;
    d->sl_val = d->recv_buf = nullptr;
    d->num_stages = 0;
    d->num_recv = 0;
    d->prepared_feat = 0;
}

extern "C" {

int gnnagg_dist_create_rank(int rank, int world, const int64_t *shard_bounds, int feat_cap, gnnagg_dist **out)
{
    if (!out || !shard_bounds || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || feat_cap < 4 || (feat_cap & 3))
        return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_create_rank: bad argument (world <= 16, feat_cap a multiple of 4)");
    for (int p = 0; p < world; ++p)
        if (shard_bounds[p + 1] < shard_bounds[p] || shard_bounds[0] != 0)
            return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_create_rank: shard_bounds must start at 0 and ascend");
    if (shard_bounds[world] > INT32_MAX) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_create_rank: source ids are int32");
    gnnagg_dist *d = new gnnagg_dist();
    d->rank = rank;
    d->world = world;
    d->pid = (int)getpid();
    d->feat_cap = feat_cap;
    memcpy(d->bounds, shard_bounds, sizeof(int64_t) * (size_t)(world + 1));
    d->rows = (int)(shard_bounds[rank + 1] - shard_bounds[rank]);
    auto fail = [&](cudaError_t e, const char *what) {
        char buf[256];
        snprintf(buf, sizeof buf, "gnnagg_dist_create_rank: %s: %s", what, cudaGetErrorString(e));
        delete d;
        return set_error(GNNAGG_ERR_CUDA, buf);
    };
    cudaError_t e = cudaGetDevice(&d->device);
    if (e != cudaSuccess) return fail(e, "cudaGetDevice");
    cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, d->device);
    const size_t shard = (((size_t)d->rows * feat_cap * sizeof(float)) + 511) & ~(size_t)511;
    d->off_x[0] = 512;  // flags occupy the first 512 bytes
    d->off_x[1] = 512 + shard;
    d->bytes = 512 + 2 * shard;
    e = cudaMalloc((void **)&d->base, d->bytes);
    if (e != cudaSuccess) return fail(e, "cudaMalloc of the peer-visible shard buffers");
    e = cudaMemset(d->base, 0, 512);
    if (e != cudaSuccess) return fail(e, "cudaMemset");
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
    e = cudaStreamCreateWithPriority(&d->comm, cudaStreamNonBlocking, hi);
    if (e != cudaSuccess) return fail(e, "cudaStreamCreateWithPriority");
    cudaEventCreateWithFlags(&d->ev_sig, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming);
    for (int s = 0; s < kMaxStages; ++s) cudaEventCreateWithFlags(&d->ev_stage[s], cudaEventDisableTiming);
    {  // load the step's kernels now: a first launch while a peer's kernels spin on this rank could wait forever
        cudaFuncAttributes attr;
        if (cudaFuncGetAttributes(&attr, halo_signal_kernel) != cudaSuccess) cudaGetLastError();
        if (cudaFuncGetAttributes(&attr, halo_done_kernel) != cudaSuccess) cudaGetLastError();
        if (cudaFuncGetAttributes(&attr, halo_pull_kernel<8>) != cudaSuccess) cudaGetLastError();
        dense_preload();
    }
    d->peer_base[rank] = d->base;
    d->peer_off_x[rank][0] = d->off_x[0];
    d->peer_off_x[rank][1] = d->off_x[1];
    d->connected = world == 1;
    *out = d;
    return GNNAGG_OK;
}

int gnnagg_dist_export(gnnagg_dist *d, void *blob)
{
    if (!d || !blob) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_export: NULL argument");
    DeviceGuard guard(d->device);
    DistBlob b;
    memset(&b, 0, sizeof b);
    b.magic = 0x676e6e61u;
    b.rank = d->rank, b.world = d->world, b.device = d->device, b.pid = d->pid;
    b.feat_cap = d->feat_cap;
    b.rows = d->rows;
    b.raw_base = (uint64_t)(uintptr_t)d->base;
    b.bytes = d->bytes;
    b.off_x[0] = d->off_x[0], b.off_x[1] = d->off_x[1];
    // a handle is only needed by OTHER processes; a failure here (IPC not permitted) is reported at connect time
    if (cudaIpcGetMemHandle(&b.mem, d->base) != cudaSuccess) {
        cudaGetLastError();
        memset(&b.mem, 0, sizeof b.mem);
    }
    memset(blob, 0, GNNAGG_DIST_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return GNNAGG_OK;
}

int gnnagg_dist_connect(gnnagg_dist *d, const void *blobs)
{
    if (!d || !blobs) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_connect: NULL argument");
    DeviceGuard guard(d->device);
    int same = 1;
    for (int p = 0; p < d->world; ++p) {
        if (p == d->rank) continue;
        DistBlob b;
        memcpy(&b, (const char *)blobs + (size_t)p * GNNAGG_DIST_BLOB_BYTES, sizeof b);
        if (b.magic != 0x676e6e61u || b.rank != p || b.world != d->world || b.feat_cap != d->feat_cap ||
            b.rows != d->bounds[p + 1] - d->bounds[p])
            return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_connect: blob does not describe the expected peer");
        if (d->peer_base[p] && d->peer_ipc[p]) cudaIpcCloseMemHandle(d->peer_base[p]);
        d->peer_base[p] = nullptr;
        d->peer_ipc[p] = false;
        if (b.pid == d->pid) {  // same process: the pointer is valid here once peer access is on
            if (b.device != d->device) {
                int can = 0;
                DT_TRY(cudaDeviceCanAccessPeer(&can, d->device, b.device));
                if (!can) return set_error(GNNAGG_ERR_CUDA, "gnnagg_dist_connect: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
                else if (e != cudaSuccess)
                    DT_TRY(e);
            }
            d->peer_base[p] = reinterpret_cast<char *>((uintptr_t)b.raw_base);
        } else {
            void *mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, b.mem, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                char buf[256];
                snprintf(buf, sizeof buf, "gnnagg_dist_connect: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
                return set_error(GNNAGG_ERR_CUDA, buf);
            }
            d->peer_base[p] = (char *)mapped;
            d->peer_ipc[p] = true;
        }
        d->peer_off_x[p][0] = b.off_x[0];
        d->peer_off_x[p][1] = b.off_x[1];
        if (b.device == d->device) ++same;
    }
    d->same_device_ranks = same;
    d->connected = true;
    return GNNAGG_OK;
}

int gnnagg_dist_create(int world, const int *devices, const int64_t *shard_bounds, int feat_cap, gnnagg_dist **out)
{
    if (!out || world < 1 || world > kMaxWorld) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_create: bad argument");
    int prev = 0;
    DT_TRY(cudaGetDevice(&prev));
    for (int r = 0; r < world; ++r) out[r] = nullptr;
    int rc = GNNAGG_OK;
    for (int r = 0; r < world && rc == GNNAGG_OK; ++r) {
        if (cudaSetDevice(devices ? devices[r] : r) != cudaSuccess) {
            rc = set_error(GNNAGG_ERR_CUDA, "gnnagg_dist_create: cudaSetDevice failed");
            break;
        }
        rc = gnnagg_dist_create_rank(r, world, shard_bounds, feat_cap, &out[r]);
    }
    if (rc == GNNAGG_OK) {
        char *blobs = new char[(size_t)world * GNNAGG_DIST_BLOB_BYTES];
        for (int r = 0; r < world && rc == GNNAGG_OK; ++r) rc = gnnagg_dist_export(out[r], blobs + (size_t)r * GNNAGG_DIST_BLOB_BYTES);
        for (int r = 0; r < world && rc == GNNAGG_OK; ++r) rc = gnnagg_dist_connect(out[r], blobs);
        delete[] blobs;
    }
    if (rc != GNNAGG_OK)
        for (int r = 0; r < world; ++r) {
            if (out[r]) gnnagg_dist_destroy(out[r]);
            out[r] = nullptr;
        }
    cudaSetDevice(prev);
    return rc;
}

int gnnagg_dist_destroy(gnnagg_dist *d)
{
    if (!d) return GNNAGG_OK;
    DeviceGuard guard(d->device);
    cudaDeviceSynchronize();
    free_graph(d);
    for (int p = 0; p < d->world; ++p)
        if (p != d->rank && d->peer_base[p] && d->peer_ipc[p]) cudaIpcCloseMemHandle(d->peer_base[p]);
    cudaFree(d->base);
    cudaFree(d->ax);
    if (d->comm) cudaStreamDestroy(d->comm);
    cudaEvent_t evs[] = {d->ev_sig, d->ev_done, d->t_m0, d->t_m1, d->t_m2, d->t_m3, d->t_c0, d->t_c1};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    for (int s = 0; s < kMaxStages; ++s)
        if (d->ev_stage[s]) cudaEventDestroy(d->ev_stage[s]);
    delete d;
    return GNNAGG_OK;
}

float *gnnagg_dist_x(gnnagg_dist *d, int buf)
{
    if (!d || buf < 0 || buf > 1) return nullptr;
    return reinterpret_cast<float *>(d->base + d->off_x[buf]);
}

int gnnagg_dist_set_graph(gnnagg_dist *d, const int *d_ptr, const int *d_idx, const float *d_val, int64_t num_e,
                          int remote_stages, void *stream)
{
    if (!d || !d_ptr || (num_e > 0 && (!d_idx || !d_val)) || num_e < 0 || num_e > INT32_MAX)
        return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_set_graph: bad argument");
    DeviceGuard guard(d->device);
    cudaStream_t st = (cudaStream_t)stream;
    free_graph(d);
    const int W = d->world, n = d->rows, m = (int)num_e;
    const int64_t total = d->bounds[W];
    const int64_t own_lo = d->bounds[d->rank], own_hi = d->bounds[d->rank + 1];
    // stages: 0 = local sources; the W-1 remote owners, taken in the order rank+1, rank+2, ... (every rank starts
    // with a different owner, so no shard is read by everybody at once), are cut into `remote_stages` groups
    int R = W > 1 ? remote_stages : 0;
    if (W > 1 && R < 1) R = 1;
    if (R > W - 1) R = W - 1;
    d->remote_stages = R;
    d->num_stages = 1 + R;
    d->stage_of[d->rank] = 0;
    for (int k = 0; k < W - 1; ++k) {
        const int p = (d->rank + 1 + k) % W;
        d->pull_order[k] = p;
        d->stage_of[p] = 1 + (int)((int64_t)k * R / (W - 1));
    }
    Bounds b;
    memset(&b, 0, sizeof b);
    b.world = W;
    for (int p = 0; p <= W; ++p) b.b[p] = d->bounds[p];
    for (int p = 0; p < W; ++p) b.stage_of[p] = d->stage_of[p];

    int *mark = nullptr, *pos = nullptr, *bad = nullptr, *idx_new = nullptr, *keys = nullptr, *item_row = nullptr;
    void *tmp = nullptr;
    int rc = GNNAGG_OK;
    auto cleanup = [&]() { cudaFree(mark), cudaFree(pos), cudaFree(bad), cudaFree(idx_new), cudaFree(keys), cudaFree(item_row), cudaFree(tmp); };
#define SG_TRY(expr)                                                                             \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _buf[512];                                                                      \
            snprintf(_buf, sizeof _buf, "gnnagg_dist_set_graph: %s: %s", #expr, cudaGetErrorString(_e)); \
            cleanup();                                                                           \
            free_graph(d);                                                                       \
            return set_error(GNNAGG_ERR_CUDA, _buf);                                             \
        }                                                                                        \
    } while (0)
    const size_t me = (size_t)(m > 0 ? m : 1);
    SG_TRY(cudaMalloc((void **)&mark, ((size_t)total + 1) * sizeof(int)));
    SG_TRY(cudaMalloc((void **)&pos, ((size_t)total + 1) * sizeof(int)));
    SG_TRY(cudaMalloc((void **)&bad, sizeof(int)));
    SG_TRY(cudaMemsetAsync(mark, 0, ((size_t)total + 1) * sizeof(int), st));
    SG_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (m > 0) dist_mark_kernel<<<nblocks(m), 256, 0, st>>>(d_idx, m, own_lo, own_hi, total, mark, bad);
    size_t need = 0;
    SG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, mark, pos, (int)(total + 1), st));
    SG_TRY(cudaMalloc(&tmp, need ? need : 1));
    SG_TRY(cub::DeviceScan::ExclusiveSum(tmp, need, mark, pos, (int)(total + 1), st));
    int h_bad = 0, h_off[kMaxWorld + 1] = {0};
    SG_TRY(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    for (int p = 0; p <= W; ++p) SG_TRY(cudaMemcpyAsync(&h_off[p], pos + d->bounds[p], sizeof(int), cudaMemcpyDeviceToHost, st));
    SG_TRY(cudaStreamSynchronize(st));
    if (h_bad) {
        cleanup();
        free_graph(d);
        return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_set_graph: a source id lies outside [0, shard_bounds[world])");
    }
    for (int p = 0; p <= W; ++p) d->recv_off[p] = h_off[p];
    d->num_recv = h_off[W];
    SG_TRY(cudaMalloc((void **)&d->recv_local, (size_t)(d->num_recv > 0 ? d->num_recv : 1) * sizeof(int)));
    SG_TRY(cudaMalloc((void **)&d->recv_buf, (size_t)(d->num_recv > 0 ? d->num_recv : 1) * d->feat_cap * sizeof(float)));
    if (total > 0) dist_fill_recv_kernel<<<nblocks(total), 256, 0, st>>>(mark, pos, total, b, d->recv_local);
    SG_TRY(cudaMalloc((void **)&idx_new, me * sizeof(int)));
    SG_TRY(cudaMalloc((void **)&keys, me * sizeof(int)));
    if (m > 0) dist_reindex_kernel<<<nblocks(m), 256, 0, st>>>(d_idx, m, pos, own_lo, own_hi, b, idx_new, keys);
    SG_TRY(cudaGetLastError());
    SG_TRY(cudaStreamSynchronize(st));
    cudaFree(mark), cudaFree(pos), cudaFree(tmp);
    mark = pos = nullptr;
    tmp = nullptr;
    // sub-CSR per stage (stable split: CSR order inside a stage), then one ordinary aggregator per stage
    int items = 0;
    rc = build_item_rows_device(d_ptr, n, m, &item_row, &items, st);
    if (rc == GNNAGG_OK)
        rc = source_slices_build_device(d_ptr, idx_new, item_row, items, n, m, d->num_stages, 1, &d->sl_ptr, &d->sl_idx, &d->sl_perm,
                                        d->sl_off, d->sl_cnt, st, keys);
    if (rc != GNNAGG_OK) {
        cleanup();
        free_graph(d);
        return rc;
    }
    const size_t padded = me + 4 * (size_t)d->num_stages;
    SG_TRY(cudaMalloc((void **)&d->sl_val, padded * sizeof(float)));
    for (int s = 0; s < d->num_stages; ++s) {
        if (d->sl_cnt[s] > 0)
            dist_gather_val_kernel<<<nblocks(d->sl_cnt[s]), 256, 0, st>>>(d_val, d->sl_perm + d->sl_off[s], d->sl_val + d->sl_off[s],
                                                                        d->sl_cnt[s]);
    }
    SG_TRY(cudaGetLastError());
    SG_TRY(cudaStreamSynchronize(st));
    cleanup();
    mark = pos = bad = idx_new = keys = item_row = nullptr;
    tmp = nullptr;
    cudaFree(d->sl_perm);  // only needed for the values
    d->sl_perm = nullptr;
    for (int s = 0; s < d->num_stages; ++s) {
        rc = gnnagg_create(d->sl_ptr + (size_t)s * ((size_t)n + 1), d->sl_idx + d->sl_off[s], nullptr, nullptr, n, d->sl_cnt[s],
                           &d->stage[s]);
        if (rc == GNNAGG_OK) rc = gnnagg_set_val(d->stage[s], d->sl_val + d->sl_off[s]);
        if (rc != GNNAGG_OK) {
            free_graph(d);
            return rc;
        }
    }
    d->num_e = num_e;
    d->launches += 6 + 3 * d->num_stages;
    return gnnagg_dist_prepare(d, d->feat_cap, stream);
#undef SG_TRY
}

int gnnagg_dist_prepare(gnnagg_dist *d, int feat, void *stream)
{
    if (!d || d->num_stages == 0) return set_error(GNNAGG_ERR_STATE, "gnnagg_dist_prepare: no graph (gnnagg_dist_set_graph)");
    if (feat < 4 || (feat & 3) || feat > d->feat_cap) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_prepare: bad feat");
    DeviceGuard guard(d->device);
    for (int s = 0; s < d->num_stages; ++s)
        if (int rc = gnnagg_prepare(d->stage[s], feat, stream)) return rc;
    // A*X of the fused layer: allocated here rather than inside the first step for the same reason
    const size_t need = (size_t)d->rows * feat;
    if (need > d->ax_cap) {
        cudaFree(d->ax);
        d->ax = nullptr;
        d->ax_cap = 0;
        DT_TRY(cudaMalloc((void **)&d->ax, (need ? need : 1) * sizeof(float)));
        d->ax_cap = need;
    }
    d->prepared_feat = feat;
    return GNNAGG_OK;
}

int gnnagg_dist_info(const gnnagg_dist *d, int64_t *num_recv, int64_t *recv_counts, int *num_stages, int64_t *stage_edges)
{
    if (!d) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_info: NULL handle");
    if (num_recv) *num_recv = d->num_recv;
    if (recv_counts)
        for (int p = 0; p < d->world; ++p) recv_counts[p] = d->recv_off[p + 1] - d->recv_off[p];
    if (num_stages) *num_stages = d->num_stages;
    if (stage_edges)
        for (int s = 0; s < d->num_stages; ++s) stage_edges[s] = d->sl_cnt[s];
    return GNNAGG_OK;
}

int gnnagg_dist_profile_enable(gnnagg_dist *d, int on)
{
    if (!d) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_profile_enable: NULL handle");
    DeviceGuard guard(d->device);
    if (on && !d->t_m0) {
        cudaEvent_t *evs[] = {&d->t_m0, &d->t_m1, &d->t_m2, &d->t_m3, &d->t_c0, &d->t_c1};
        for (cudaEvent_t *e : evs) DT_TRY(cudaEventCreate(e));
    }
    d->prof = on != 0;
    return GNNAGG_OK;
}

int gnnagg_dist_profile_read(gnnagg_dist *d, float *ms)
{
    if (!d || !ms || !d->t_m0) return set_error(GNNAGG_ERR_STATE, "gnnagg_dist_profile_read: profiling not enabled");
    DeviceGuard guard(d->device);
    DT_TRY(cudaEventSynchronize(d->t_m3));
    DT_TRY(cudaEventSynchronize(d->t_c1));
    DT_TRY(cudaEventElapsedTime(&ms[0], d->t_c0, d->t_c1));  // halo exchange: first pull issued .. last pull landed
    DT_TRY(cudaEventElapsedTime(&ms[1], d->t_m0, d->t_m3));  // whole step
    DT_TRY(cudaEventElapsedTime(&ms[2], d->t_m0, d->t_m1));  // stage 0 (local sources)
    DT_TRY(cudaEventElapsedTime(&ms[3], d->t_m2, d->t_m3));  // dense combination
    return GNNAGG_OK;
}

int gnnagg_dist_check(gnnagg_dist *d)
{
    if (!d) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_check: NULL handle");
    DeviceGuard guard(d->device);
    uint32_t err = 0;
    DT_TRY(cudaMemcpy(&err, &flags_of(d->base)->err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        char buf[160];
        snprintf(buf, sizeof buf, "halo protocol: rank %d gave up waiting for rank %u (%s flag, code 0x%x)", d->rank, err & 0xffu,
                 (err & 0x100u) ? "done" : "ready", err);
        return set_error(GNNAGG_ERR_STATE, buf);
    }
    return GNNAGG_OK;
}

int64_t gnnagg_dist_launch_count(const gnnagg_dist *d)
{
    if (!d) return 0;
    int64_t total = d->launches;
    for (int s = 0; s < d->num_stages; ++s) total += gnnagg_launch_count(d->stage[s]);
    return total;
}

// One step: Y = A_block * X with X = the shards `buf` of all ranks.  flags: GNNAGG_DIST_NO_EXCHANGE re-uses the
// receive buffer of the previous step (kernels-only timing).
static int dist_run(gnnagg_dist *d, int buf, float *Y, const float *W, float *H, int feat_in, int feat_out, int flags,
                    cudaStream_t st)
{
    if (!d || (!Y && d->rows > 0) || buf < 0 || buf > 1) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_gcn_run: bad argument");
    if (!d->connected) return set_error(GNNAGG_ERR_STATE, "gnnagg_dist_gcn_run: peers not connected (gnnagg_dist_connect)");
    if (d->num_stages == 0) return set_error(GNNAGG_ERR_STATE, "gnnagg_dist_gcn_run: no graph (gnnagg_dist_set_graph)");
    if (feat_in < 4 || (feat_in & 3) || feat_in > d->feat_cap)
        return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_gcn_run: feat must be a multiple of 4, at most feat_cap");
    DeviceGuard guard(d->device);
    const int Wd = d->world;
    const bool exchange = Wd > 1 && !(flags & GNNAGG_DIST_NO_EXCHANGE);
    // per-stage fix-up tables for this feature width (set_graph builds them for feat_cap).  Building them waits for the
    // device, which is harmless with one process per GPU; a process that drives SEVERAL ranks must call
    // gnnagg_dist_prepare for every rank before the first step of a new width (a device-wide wait while another rank's
    // kernels spin on this rank's signal would never return)
    if (d->prepared_feat != feat_in)
        if (int rc = gnnagg_dist_prepare(d, feat_in, st)) return rc;
    const float *xs = reinterpret_cast<const float *>(d->base + d->off_x[buf]);
    if (d->prof) DT_TRY(cudaEventRecord(d->t_m0, st));
    if (exchange) {
        const uint32_t epoch = ++d->epoch;
        PeerTable peers;
        memset(&peers, 0, sizeof peers);
        for (int p = 0; p < Wd; ++p) peers.flags[p] = flags_of(d->peer_base[p]);
        HaloFlags *mine = flags_of(d->base);
        halo_signal_kernel<<<1, 32, 0, st>>>(mine, peers, d->rank, Wd, epoch);
        DT_TRY(cudaPeekAtLastError());
        DT_TRY(cudaEventRecord(d->ev_sig, st));
        DT_TRY(cudaStreamWaitEvent(d->comm, d->ev_sig, 0));
        if (d->prof) DT_TRY(cudaEventRecord(d->t_c0, d->comm));
        const int F4 = feat_in / 4;
        int shift = -1;
        for (int s = 0; s < 16; ++s)
            if ((1 << s) == F4) shift = s;
        int stage_now = 1;
        for (int k = 0; k < Wd - 1; ++k) {
            const int p = d->pull_order[k];
            if (d->stage_of[p] != stage_now) {  // the previous stage is complete
                DT_TRY(cudaEventRecord(d->ev_stage[stage_now], d->comm));
                stage_now = d->stage_of[p];
            }
            const int64_t cnt = d->recv_off[p + 1] - d->recv_off[p];
            if (cnt > 0) {
                const int64_t count4 = cnt * F4;
                // 2 CTAs per SM saturate the link; ranks sharing one device (tests) split that, so that CTAs spinning on
                // a peer's flag can never fill the device and keep that peer's signal kernel out
                int64_t grid = (count4 + 256 * 8 - 1) / (256 * 8);
                const int64_t cap = d->same_device_ranks > 1 ? std::max(4, d->sm_count / (2 * d->same_device_ranks)) : 2 * d->sm_count;
                if (grid > cap) grid = cap;
                halo_pull_kernel<8><<<(unsigned)grid, 256, 0, d->comm>>>(
                    reinterpret_cast<const float4 *>(d->peer_base[p] + d->peer_off_x[p][buf]), d->recv_local + d->recv_off[p],
                    reinterpret_cast<float4 *>(d->recv_buf + (size_t)d->recv_off[p] * feat_in), count4, F4, shift, &mine->ready[p], epoch,
                    &mine->err, (uint32_t)p);
                DT_TRY(cudaPeekAtLastError());
                ++d->launches;
            }
        }
        DT_TRY(cudaEventRecord(d->ev_stage[stage_now], d->comm));
        halo_done_kernel<<<1, 32, 0, d->comm>>>(peers, d->rank, Wd, epoch);
        DT_TRY(cudaPeekAtLastError());
        DT_TRY(cudaEventRecord(d->ev_done, d->comm));
        if (d->prof) DT_TRY(cudaEventRecord(d->t_c1, d->comm));
        d->launches += 2;
    } else if (d->prof) {
        DT_TRY(cudaEventRecord(d->t_c0, st));
        DT_TRY(cudaEventRecord(d->t_c1, st));
    }
    float *agg_out = Y;
    if (W) {
        const size_t need = (size_t)d->rows * feat_in;
        if (need > d->ax_cap) {
            cudaFree(d->ax);
            d->ax = nullptr;
            d->ax_cap = 0;
            DT_TRY(cudaMalloc((void **)&d->ax, (need ? need : 1) * sizeof(float)));
            d->ax_cap = need;
        }
        agg_out = d->ax;
    }
    // stage 0: sources of the own shard, no communication needed
    if (int rc = gnnagg_gcn_run_acc(d->stage[0], xs, agg_out, feat_in, 0, st)) return rc;
    if (d->prof) DT_TRY(cudaEventRecord(d->t_m1, st));
    for (int s = 1; s < d->num_stages; ++s) {
        if (exchange) DT_TRY(cudaStreamWaitEvent(st, d->ev_stage[s], 0));
        if (int rc = gnnagg_gcn_run_acc(d->stage[s], d->recv_buf, agg_out, feat_in, 1, st)) return rc;
    }
    if (d->prof) DT_TRY(cudaEventRecord(d->t_m2, st));
    if (W && d->rows > 0) {
        if (int rc = gnnagg_dense_nn(d->ax, W, H, d->rows, feat_out, feat_in, st)) return rc;
        d->launches += 2;
    }
    if (exchange) DT_TRY(cudaStreamWaitEvent(st, d->ev_done, 0));  // the caller's stream covers the comm stream's work too
    if (d->prof) DT_TRY(cudaEventRecord(d->t_m3, st));
    return GNNAGG_OK;
}

int gnnagg_dist_gcn_run(gnnagg_dist *d, int buf, float *Y, int feat, int flags, void *stream)
{
    return dist_run(d, buf, Y, nullptr, nullptr, feat, feat, flags, (cudaStream_t)stream);
}

int gnnagg_dist_gcn_layer(gnnagg_dist *d, int buf, const float *W, float *H, int feat_in, int feat_out, int flags, void *stream)
{
    if (!W || (!H && d && d->rows > 0)) return set_error(GNNAGG_ERR_ARG, "gnnagg_dist_gcn_layer: NULL argument");
    return dist_run(d, buf, H, W, H, feat_in, feat_out, flags, (cudaStream_t)stream);
}

}  // extern "C"
