// lsh_reorder.cpp -- locality-aware vertex reordering: the preprocessing pass that produces a
// <dset>.reorder<suffix> permutation (consumed by gnnagg_graph_load / load_graph).
//
// Deterministic re-statement of the reference's offline script script/cluster2.py:
//   MinHash(64 perms) + LSH(threshold 0.2) candidate pairs (:29-37, :80-96)
//   -> exact Jaccard of the two neighbour lists (:44-49)
//   -> max-heap greedy union-find clustering, cluster size cap 64 with "deleted" freeze, non-root
//      pairs re-queued as root pairs scored on the ROOTS' OWN lists (:108-153)
//   -> clusters emitted in order of first member, members ascending (:156-171).
// cluster2.py takes MinHash/LSH from datasketch (not available, and not reproducible: its query
// order depends on PYTHONHASHSEED and heap ties on insertion history), so the hash family and the
// tie-breaks are specified here (oracle/cluster2_port.py is the executable spec this file is
// tested against): splitmix64 vertex hash, 2^61-1 affine permutations, b bands x r rows (28 x 2 for
// threshold 0.2), bucket window of 32 neighbours by id, heap key (similarity desc, min id, max id).
// Not a port of the script: signatures are computed in parallel over vertices, buckets come from
// one sort per band instead of Python dicts, Jaccard is a sorted-list merge, the pair set is a
// flat hash set of 64-bit keys.
#include <algorithm>
#include <cstdint>
#include <queue>
#include <unordered_set>
#include <vector>

#include "gnnagg.h"
#include "internal.h"

namespace {

constexpr uint64_t kP61 = (1ull << 61) - 1;
constexpr uint32_t kMaxHash = 0xFFFFFFFFu;
constexpr int kWindow = 32;

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Pair {
    double sim;
    int lo, hi, p1, p2;
};
struct PairOrder {  // priority_queue pops the LARGEST: higher similarity, then smaller (lo, hi)
    bool operator()(const Pair &a, const Pair &b) const
    {
        if (a.sim != b.sim) return a.sim < b.sim;
        if (a.lo != b.lo) return a.lo > b.lo;
        return a.hi > b.hi;
    }
};

struct Reorderer {
    const int *ptr, *idx;
    int numv;
    std::vector<int> uptr, uidx;  // per-vertex sorted unique neighbour lists (set semantics of jd())

    void build_unique()
    {
        uptr.assign((size_t)numv + 1, 0);
        uidx.resize((size_t)ptr[numv]);
        std::vector<int> tmp;
        int64_t pos = 0;
        for (int i = 0; i < numv; ++i) {
            tmp.assign(idx + ptr[i], idx + ptr[i + 1]);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            std::copy(tmp.begin(), tmp.end(), uidx.begin() + pos);
            pos += (int64_t)tmp.size();
            uptr[i + 1] = (int)pos;
        }
    }
    double jaccard(int a, int b) const
    {
        const int *x = uidx.data() + uptr[a], *xe = uidx.data() + uptr[a + 1];
        const int *y = uidx.data() + uptr[b], *ye = uidx.data() + uptr[b + 1];
        if (x == xe || y == ye) return 0.0;
        const int64_t la = xe - x, lb = ye - y;
        int64_t inter = 0;
        while (x < xe && y < ye) {
            if (*x < *y)
                ++x;
            else if (*y < *x)
                ++y;
            else
                ++inter, ++x, ++y;
        }
        return (double)inter / (double)(la + lb - inter);
    }
};

}  // namespace

extern "C" int gnnagg_lsh_reorder(const int *ptr, const int *idx, int num_v, int num_e, int num_perm, int bands,
                                  int rows_per_band, int cluster_cap, uint64_t seed, int *rows)
{
    using gnnagg::set_error;
    if (!ptr || (!idx && num_e > 0) || !rows || num_v < 0) return set_error(GNNAGG_ERR_ARG, "gnnagg_lsh_reorder: bad argument");
    if (num_perm <= 0) num_perm = 64;  // script/cluster2.py:6
    if (bands == 0) bands = 28, rows_per_band = 2;  // datasketch optimum for threshold 0.2 (:7), SURVEY 8(c)
    if (cluster_cap <= 0) cluster_cap = 64;         // :10
    if (bands > 0 && (rows_per_band <= 0 || (int64_t)bands * rows_per_band > num_perm))
        return set_error(GNNAGG_ERR_ARG, "gnnagg_lsh_reorder: bands*rows_per_band must be <= num_perm");
    const int numv = num_v;
    Reorderer R{ptr, idx, numv, {}, {}};
    R.build_unique();

    // ---- candidate pairs -------------------------------------------------------------------
    std::vector<std::vector<int>> cand((size_t)numv);
    if (bands > 0) {
        std::vector<uint64_t> pa((size_t)num_perm), pb((size_t)num_perm);
        for (int k = 0; k < num_perm; ++k) {
            pa[k] = 1 + splitmix64((seed << 32) + 2 * (uint64_t)k) % (kP61 - 1);
            pb[k] = splitmix64((seed << 32) + 2 * (uint64_t)k + 1) % kP61;
        }
        const int used = bands * rows_per_band;
        std::vector<uint32_t> sig((size_t)numv * used);
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < numv; ++i) {
            uint32_t *s = sig.data() + (size_t)i * used;
            for (int k = 0; k < used; ++k) s[k] = kMaxHash;
            for (int e = R.uptr[i]; e < R.uptr[i + 1]; ++e) {
                const uint64_t h = splitmix64((uint64_t)R.uidx[e]) & kMaxHash;
                for (int k = 0; k < used; ++k) {
                    const uint32_t v = (uint32_t)((((unsigned __int128)pa[k] * h + pb[k]) % kP61) & kMaxHash);
                    if (v < s[k]) s[k] = v;
                }
            }
        }
        std::vector<int> order((size_t)numv);
        for (int j = 0; j < bands; ++j) {
            for (int i = 0; i < numv; ++i) order[i] = i;
            const uint32_t *base = sig.data() + (size_t)j * rows_per_band;
            auto key_less = [&](int a, int b) {
                const uint32_t *x = base + (size_t)a * used, *y = base + (size_t)b * used;
                for (int r = 0; r < rows_per_band; ++r)
                    if (x[r] != y[r]) return x[r] < y[r];
                return a < b;  // members of a bucket ascending by id
            };
            auto key_eq = [&](int a, int b) {
                const uint32_t *x = base + (size_t)a * used, *y = base + (size_t)b * used;
                for (int r = 0; r < rows_per_band; ++r)
                    if (x[r] != y[r]) return false;
                return true;
            };
            std::sort(order.begin(), order.end(), key_less);
            for (int s = 0; s < numv;) {
                int e = s + 1;
                while (e < numv && key_eq(order[s], order[e])) ++e;
                for (int p = s; p < e; ++p)
                    for (int q = std::max(s, p - kWindow); q < std::min(e, p + kWindow + 1); ++q)
                        if (q != p) cand[order[p]].push_back(order[q]);
                s = e;
            }
        }
    } else {  // exhaustive: every pair sharing a neighbour (test mode, small graphs)
        std::vector<std::vector<int>> owners;
        int max_id = -1;
        for (int e = 0; e < (int)R.uidx.size(); ++e) max_id = std::max(max_id, R.uidx[e]);
        owners.resize((size_t)max_id + 1);
        for (int i = 0; i < numv; ++i)
            for (int e = R.uptr[i]; e < R.uptr[i + 1]; ++e) owners[R.uidx[e]].push_back(i);
        for (auto &members : owners)
            for (int a : members)
                for (int b : members)
                    if (a != b) cand[a].push_back(b);
    }
    for (auto &c : cand) {
        std::sort(c.begin(), c.end());
        c.erase(std::unique(c.begin(), c.end()), c.end());
    }

    // ---- heap of scored pairs ----------------------------------------------------------------
    auto makenum = [numv](int a, int b) -> uint64_t {
        return a <= b ? (uint64_t)a * (uint64_t)numv + (uint64_t)b : (uint64_t)b * (uint64_t)numv + (uint64_t)a;
    };
    std::priority_queue<Pair, std::vector<Pair>, PairOrder> heap;
    std::unordered_set<uint64_t> sset;
    auto put = [&](int p1, int p2) {
        heap.push(Pair{R.jaccard(p1, p2), std::min(p1, p2), std::max(p1, p2), p1, p2});
        sset.insert(makenum(p1, p2));
    };
    for (int i = 0; i < numv; ++i) {
        if (ptr[i] == ptr[i + 1]) continue;
        for (int c : cand[i]) {
            if (c == i || sset.count(makenum(i, c))) continue;
            put(i, c);
        }
        std::vector<int>().swap(cand[i]);
    }

    // ---- greedy size-capped union-find ---------------------------------------------------------
    std::vector<int> cluster_id((size_t)numv), cluster_sz((size_t)numv, 1);
    std::vector<char> deleted((size_t)numv, 0);
    for (int i = 0; i < numv; ++i) cluster_id[i] = i;
    auto root = [&](int i) {
        while (i != cluster_id[i]) {
            cluster_id[i] = cluster_id[cluster_id[i]];
            i = cluster_id[i];
        }
        return i;
    };
    int64_t num_cluster = numv;
    while (!heap.empty() && num_cluster > 0) {
        const Pair top = heap.top();
        heap.pop();
        int p1 = top.p1, p2 = top.p2;
        sset.erase(makenum(p1, p2));
        if (p1 == cluster_id[p1] && p2 == cluster_id[p2]) {
            if (deleted[p1] || deleted[p2]) continue;
            // the smaller cluster joins the larger one; on a tie p2 joins p1
            const int into = cluster_sz[p1] < cluster_sz[p2] ? p2 : p1, from = into == p1 ? p2 : p1;
            cluster_id[from] = into;
            --num_cluster;
            cluster_sz[into] += cluster_sz[from];
            if (cluster_sz[into] >= cluster_cap) {
                deleted[into] = 1;
                --num_cluster;
            }
        } else {
            p1 = root(p1);
            p2 = root(p2);
            if (deleted[p1] || deleted[p2]) continue;
            if (p1 != p2 && !sset.count(makenum(p1, p2))) put(p1, p2);
        }
    }

    // ---- emit: clusters by first member, members ascending -----------------------------------
    std::vector<int> slot((size_t)numv, -1), count;
    std::vector<int> root_of((size_t)numv);
    int nclusters = 0;
    for (int i = 0; i < numv; ++i) {
        const int r = root(i);
        root_of[i] = r;
        if (slot[r] < 0) {
            slot[r] = nclusters++;
            count.push_back(0);
        }
        ++count[slot[r]];
    }
    std::vector<int64_t> start((size_t)nclusters + 1, 0);
    for (int c = 0; c < nclusters; ++c) start[c + 1] = start[c] + count[c];
    for (int i = 0; i < numv; ++i) rows[start[slot[root_of[i]]]++] = i;
    return GNNAGG_OK;
}
