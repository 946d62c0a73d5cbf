// dist_check.cu -- a plain C++ caller of the multi-GPU entry points of include/gnnagg.h (gnnagg_dist_*), the
// path a maintainer of the reference would take from its multi-GPU leftovers (GPUNUM / gptrs / gidxs,
// include/util.h:39-57; syncAll, :135-142).  ONE process drives `world` ranks: one per visible GPU, or -- on a
// single-GPU box -- several ranks on device 0 (peer pointers are then ordinary device pointers; the flag
// protocol, the pull kernels and the staged accumulation are exercised all the same).
// A random power-law-ish graph is partitioned by destination row; the distributed result is compared with
//   (a) a scalar fp64 CPU CSR SpMM (gate: |y - y64| <= 1e-5 * sum|terms|), and
//   (b) the single-GPU gnnagg_gcn_run on device 0 (same gate),
// for the aggregation alone and for the fused layer, for two consecutive steps with different X (epoch flags).
// Prints "DIST_CHECK ok" and exits 0 when everything agrees.
//   usage: dist_check [world] [rows_per_rank] [avg_degree] [feat] [remote_stages]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "gnnagg.h"

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)
#define GK(expr)                                                                              \
    do {                                                                                      \
        int _rc = (expr);                                                                     \
        if (_rc != GNNAGG_OK) {                                                               \
            fprintf(stderr, "gnnagg error %d (%s) at %s:%d\n", _rc, gnnagg_last_error(), __FILE__, __LINE__); \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

template <class T>
static T *to_device(const std::vector<T> &v)
{
    T *p = nullptr;
    cudaMalloc((void **)&p, (v.size() ? v.size() : 1) * sizeof(T));
    if (!v.empty()) cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}

int main(int argc, char **argv)
{
    int ngpu = 0;
    CK(cudaGetDeviceCount(&ngpu));
    if (ngpu < 1) {
        fprintf(stderr, "dist_check needs a GPU (there is no CPU path)\n");
        return 2;
    }
    int world = argc > 1 ? atoi(argv[1]) : (ngpu > 1 ? (ngpu > 8 ? 8 : ngpu) : 3);
    const int rows_per = argc > 2 ? atoi(argv[2]) : 3000;
    const double avg = argc > 3 ? atof(argv[3]) : 24.0;
    const int F = argc > 4 ? atoi(argv[4]) : 64;
    const int stages = argc > 5 ? atoi(argv[5]) : 2;
    if (world < 1 || world > GNNAGG_DIST_MAX_WORLD) return 2;
    std::vector<int> devices(world);
    for (int r = 0; r < world; ++r) devices[r] = r % ngpu;

    // ---- graph: ragged shards (rank r owns rows_per + 17 r rows), a hub row, empty rows, duplicate edges
    std::vector<int64_t> bounds(world + 1, 0);
    for (int r = 0; r < world; ++r) bounds[r + 1] = bounds[r] + rows_per + 17 * r;
    const int n = (int)bounds[world];
    std::mt19937_64 rng(123);
    std::vector<int> ptr(n + 1, 0), idx;
    std::vector<float> val;
    std::poisson_distribution<int> pdeg(avg);
    std::uniform_real_distribution<float> uval(0.1f, 1.1f);
    for (int v = 0; v < n; ++v) {
        int deg = (rng() % 5 == 0) ? 0 : pdeg(rng);
        if (v == n / 3) deg = 20000;  // spans many items: carry fix-up inside every stage
        for (int k = 0; k < deg; ++k) {
            // half of the sources come from a small hot set (re-referenced rows), the rest uniformly
            const int u = (rng() & 1) ? (int)(rng() % (uint64_t)(n / 50 + 1)) * 50 % n : (int)(rng() % (uint64_t)n);
            idx.push_back(u);
            val.push_back(uval(rng));
        }
        ptr[v + 1] = (int)idx.size();
    }
    const int m = (int)idx.size();
    const int OUT = F;
    std::vector<float> W((size_t)F * OUT);
    std::normal_distribution<float> nrm(0.f, 1.f);
    for (auto &w : W) w = nrm(rng) / std::sqrt((float)F);

    gnnagg_dist *ranks[GNNAGG_DIST_MAX_WORLD] = {nullptr};
    GK(gnnagg_dist_create(world, devices.data(), bounds.data(), F, ranks));
    std::vector<int *> d_ptr(world), d_idx(world);
    std::vector<float *> d_val(world), d_y(world), d_h(world), d_w(world);
    // one stream per rank: ranks that share a device must not share a stream (a rank's stage waits for rows whose
    // owner has to get its "ready" signal out -- on a common stream that signal would be queued behind the wait)
    std::vector<cudaStream_t> streams(world);
    for (int r = 0; r < world; ++r) {
        CK(cudaSetDevice(devices[r]));
        CK(cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking));
        const int r0 = (int)bounds[r], r1 = (int)bounds[r + 1], e0 = ptr[r0], e1 = ptr[r1];
        std::vector<int> lp(r1 - r0 + 1);
        for (int i = 0; i <= r1 - r0; ++i) lp[i] = ptr[r0 + i] - e0;
        d_ptr[r] = to_device(lp);
        d_idx[r] = to_device(std::vector<int>(idx.begin() + e0, idx.begin() + e1));
        d_val[r] = to_device(std::vector<float>(val.begin() + e0, val.begin() + e1));
        d_w[r] = to_device(W);
        CK(cudaMalloc((void **)&d_y[r], (size_t)(r1 - r0 + 1) * F * sizeof(float)));
        CK(cudaMalloc((void **)&d_h[r], (size_t)(r1 - r0 + 1) * OUT * sizeof(float)));
        GK(gnnagg_dist_set_graph(ranks[r], d_ptr[r], d_idx[r], d_val[r], e1 - e0, stages, nullptr));
    }
    GK(gnnagg_dist_connect_local(ranks, world));
    {
        int64_t nrecv = 0, counts[GNNAGG_DIST_MAX_WORLD] = {0}, edges[GNNAGG_DIST_MAX_WORLD] = {0};
        int nst = 0;
        GK(gnnagg_dist_info(ranks[0], &nrecv, counts, &nst, edges, nullptr));
        printf("world %d on %d GPU(s): n=%d m=%d F=%d; rank 0 receives %lld remote rows in %d stages\n", world, ngpu, n, m, F,
               (long long)nrecv, nst);
    }

    // single-GPU comparison run on device 0
    CK(cudaSetDevice(0));
    int *g_ptr = to_device(ptr), *g_idx = to_device(idx);
    float *g_val = to_device(val), *g_x = nullptr, *g_y = nullptr;
    CK(cudaMalloc((void **)&g_x, (size_t)n * F * sizeof(float)));
    CK(cudaMalloc((void **)&g_y, (size_t)n * F * sizeof(float)));
    gnnagg_aggregator *single = nullptr;
    GK(gnnagg_create(g_ptr, g_idx, nullptr, nullptr, n, m, &single));
    GK(gnnagg_set_val(single, g_val));

    int failures = 0;
    for (int step = 0; step < 2; ++step) {
        const int buf = step & 1;  // alternate the two peer-visible shard buffers
        std::vector<float> X((size_t)n * F);
        for (auto &x : X) x = nrm(rng);
        // fp64 CPU reference + scale
        std::vector<double> y64((size_t)n * F, 0.0), sc((size_t)n * F, 0.0);
        for (int v = 0; v < n; ++v)
            for (int e = ptr[v]; e < ptr[v + 1]; ++e)
                for (int c = 0; c < F; ++c) {
                    const double t = (double)val[e] * (double)X[(size_t)idx[e] * F + c];
                    y64[(size_t)v * F + c] += t;
                    sc[(size_t)v * F + c] += std::fabs(t);
                }
        for (int r = 0; r < world; ++r) {
            CK(cudaSetDevice(devices[r]));
            CK(cudaMemcpy(gnnagg_dist_x(ranks[r], buf), X.data() + (size_t)bounds[r] * F,
                          (size_t)(bounds[r + 1] - bounds[r]) * F * sizeof(float), cudaMemcpyHostToDevice));
            CK(cudaDeviceSynchronize());
        }
        // all ranks are launched from this one thread; nothing below blocks the host until the final synchronisation
        for (int r = 0; r < world; ++r) {
            CK(cudaSetDevice(devices[r]));
            GK(gnnagg_dist_gcn_run(ranks[r], buf, d_y[r], F, 0, streams[r]));
            GK(gnnagg_dist_gcn_layer(ranks[r], buf, d_w[r], d_h[r], F, OUT, 0, streams[r]));
        }
        std::vector<float> Y((size_t)n * F), H((size_t)n * OUT);
        for (int r = 0; r < world; ++r) {
            CK(cudaSetDevice(devices[r]));
            CK(cudaDeviceSynchronize());
            GK(gnnagg_dist_check(ranks[r]));
            CK(cudaMemcpy(Y.data() + (size_t)bounds[r] * F, d_y[r], (size_t)(bounds[r + 1] - bounds[r]) * F * sizeof(float),
                          cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(H.data() + (size_t)bounds[r] * OUT, d_h[r], (size_t)(bounds[r + 1] - bounds[r]) * OUT * sizeof(float),
                          cudaMemcpyDeviceToHost));
        }
        CK(cudaSetDevice(0));
        CK(cudaMemcpy(g_x, X.data(), X.size() * sizeof(float), cudaMemcpyHostToDevice));
        GK(gnnagg_gcn_run(single, g_x, g_y, F, 0, nullptr));
        std::vector<float> Y1((size_t)n * F);
        CK(cudaMemcpy(Y1.data(), g_y, Y1.size() * sizeof(float), cudaMemcpyDeviceToHost));
        double worst = 0.0, worst1 = 0.0, worst_h = 0.0;
        for (size_t i = 0; i < Y.size(); ++i) {
            const double bound = 1e-5 * sc[i] + 1e-30;
            worst = std::fmax(worst, std::fabs((double)Y[i] - y64[i]) / bound);
            worst1 = std::fmax(worst1, std::fabs((double)Y1[i] - y64[i]) / bound);
        }
        for (int v = 0; v < n; ++v)
            for (int o = 0; o < OUT; ++o) {
                double h = 0.0, hs = 0.0;
                for (int k = 0; k < F; ++k) {
                    h += y64[(size_t)v * F + k] * (double)W[(size_t)k * OUT + o];
                    hs += sc[(size_t)v * F + k] * std::fabs((double)W[(size_t)k * OUT + o]);
                }
                worst_h = std::fmax(worst_h, std::fabs((double)H[(size_t)v * OUT + o] - h) / (1e-5 * hs + 1e-30));
            }
        printf("step %d (shard buffer %d): worst err/bound  dist Y %.3f  single-GPU Y %.3f  dist H %.3f\n", step, buf, worst, worst1,
               worst_h);
        failures += !(worst <= 1.0) + !(worst1 <= 1.0) + !(worst_h <= 1.0);
    }
    for (int r = 0; r < world; ++r) GK(gnnagg_dist_disconnect(ranks[r]));
    for (int r = 0; r < world; ++r) GK(gnnagg_dist_destroy(ranks[r]));
    gnnagg_destroy(single);
    if (failures) {
        printf("DIST_CHECK FAILED (%d)\n", failures);
        return 1;
    }
    printf("DIST_CHECK ok\n");
    return 0;
}
