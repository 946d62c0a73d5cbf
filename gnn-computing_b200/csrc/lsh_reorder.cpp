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
#include <parallel/algorithm>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <queue>
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

// open-addressing set of 64-bit keys (the "pairs currently queued" set of cluster2.py:70,96,126);
// linear probing, tombstones, grows at 70 % load.  Key 0 is stored shifted by one.
class KeySet {
public:
    explicit KeySet(size_t expected) { rehash(expected * 2 + 16); }
    bool contains(uint64_t k) const
    {
        k += 2;
        for (size_t i = slot(k);; i = (i + 1) & mask_) {
            if (tab_[i] == k) return true;
            if (tab_[i] == kEmpty) return false;
        }
    }
    void insert(uint64_t k)
    {
        if ((used_ + 1) * 10 > (mask_ + 1) * 7) rehash((live_ + 1) * 4);
        k += 2;
        size_t i = slot(k), grave = SIZE_MAX;
        for (;; i = (i + 1) & mask_) {
            if (tab_[i] == k) return;
            if (tab_[i] == kTomb && grave == SIZE_MAX) grave = i;
            if (tab_[i] == kEmpty) break;
        }
        if (grave != SIZE_MAX)
            i = grave;
        else
            ++used_;
        tab_[i] = k;
        ++live_;
    }
    void erase(uint64_t k)
    {
        k += 2;
        for (size_t i = slot(k);; i = (i + 1) & mask_) {
            if (tab_[i] == k) {
                tab_[i] = kTomb;
                --live_;
                return;
            }
            if (tab_[i] == kEmpty) return;
        }
    }

private:
    static constexpr uint64_t kEmpty = 0, kTomb = 1;
    size_t slot(uint64_t k) const { return (size_t)(splitmix64(k)) & mask_; }
    void rehash(size_t want)
    {
        size_t cap = 16;
        while (cap < want) cap <<= 1;
        std::vector<uint64_t> old;
        old.swap(tab_);
        tab_.assign(cap, kEmpty);
        mask_ = cap - 1;
        used_ = live_ = 0;
        for (uint64_t k : old)
            if (k > kTomb) {
                size_t i = slot(k);
                while (tab_[i] != kEmpty) i = (i + 1) & mask_;
                tab_[i] = k;
                ++used_, ++live_;
            }
    }
    std::vector<uint64_t> tab_;
    size_t mask_ = 0, used_ = 0, live_ = 0;
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
    const bool trace = getenv("GNNAGG_REORDER_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[lsh_reorder] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    Reorderer R{ptr, idx, numv, {}, {}};
    R.build_unique();
    lap("unique neighbour lists");

    // ---- candidate pairs: unordered (lo,hi) keys, deduplicated by one sort -------------------------
    // Candidates are symmetric and vertices are visited in ascending id (cluster2.py:80-96), so every
    // pair is first met from its smaller endpoint: Pair(p1 = lo, p2 = hi).  Pairs that involve an
    // empty row can only be met from the non-empty endpoint; they score 0 either way.
    std::vector<uint64_t> keys;
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
        lap("minhash signatures");
        std::vector<std::vector<uint64_t>> band_keys((size_t)bands);
#pragma omp parallel for schedule(dynamic, 1)
        for (int j = 0; j < bands; ++j) {
            // one 64-bit bucket key per vertex (hash of the band's rows; equality re-checked below)
            std::vector<std::pair<uint64_t, int>> order((size_t)numv);
            const uint32_t *base = sig.data() + (size_t)j * rows_per_band;
            for (int i = 0; i < numv; ++i) {
                const uint32_t *x = base + (size_t)i * used;
                uint64_t h = 0x243F6A8885A308D3ull;
                for (int r = 0; r < rows_per_band; ++r) h = splitmix64(h ^ x[r]);
                order[i] = {h, i};
            }
            std::sort(order.begin(), order.end());  // bucket members end up ascending by id
            auto same = [&](int a, int b) {
                const uint32_t *x = base + (size_t)a * used, *y = base + (size_t)b * used;
                for (int r = 0; r < rows_per_band; ++r)
                    if (x[r] != y[r]) return false;
                return true;
            };
            std::vector<uint64_t> &out = band_keys[j];
            for (int s0 = 0; s0 < numv;) {
                int e0 = s0 + 1;
                while (e0 < numv && order[e0].first == order[s0].first && same(order[s0].second, order[e0].second)) ++e0;
                for (int p = s0; p < e0; ++p)
                    for (int q = p + 1; q < std::min(e0, p + kWindow + 1); ++q) {
                        const int a = order[p].second, b = order[q].second;  // a < b
                        if (ptr[a] == ptr[a + 1] && ptr[b] == ptr[b + 1]) continue;  // two empty rows: never queried
                        out.push_back(((uint64_t)a << 32) | (uint32_t)b);
                    }
                s0 = e0;
            }
        }
        size_t total = 0;
        for (auto &v : band_keys) total += v.size();
        keys.reserve(total);
        for (auto &v : band_keys) {
            keys.insert(keys.end(), v.begin(), v.end());
            std::vector<uint64_t>().swap(v);
        }
    } else {  // exhaustive: every pair sharing a neighbour (test mode, small graphs)
        std::vector<std::vector<int>> owners;
        int max_id = -1;
        for (int e = 0; e < (int)R.uidx.size(); ++e) max_id = std::max(max_id, R.uidx[e]);
        owners.resize((size_t)max_id + 1);
        for (int i = 0; i < numv; ++i)
            for (int e = R.uptr[i]; e < R.uptr[i + 1]; ++e) owners[R.uidx[e]].push_back(i);
        for (auto &members : owners)
            for (size_t x = 0; x < members.size(); ++x)
                for (size_t y = x + 1; y < members.size(); ++y)
                    keys.push_back(((uint64_t)members[x] << 32) | (uint32_t)members[y]);
    }
    lap("band buckets");
    __gnu_parallel::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    lap("candidate dedup");

    // ---- scored pairs -> heap -----------------------------------------------------------------
    auto makenum = [numv](int a, int b) -> uint64_t {
        return a <= b ? (uint64_t)a * (uint64_t)numv + (uint64_t)b : (uint64_t)b * (uint64_t)numv + (uint64_t)a;
    };
    // Initial pairs carry their index into `keys` in p1 and p2 = -1 until they are popped (their endpoints are lo/hi).
    const size_t npairs = keys.size();
    if (npairs > (size_t)INT32_MAX) return set_error(GNNAGG_ERR_ARG, "gnnagg_lsh_reorder: more than 2^31 candidate pairs");
    std::vector<Pair> scored(npairs);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t k = 0; k < (int64_t)npairs; ++k) {
        const int a = (int)(keys[k] >> 32), b = (int)(keys[k] & 0xFFFFFFFFu);
        scored[k] = Pair{R.jaccard(a, b), a, b, (int)k, -1};
    }
    // The queue of cluster2.py is one max-heap plus a set of the pairs currently queued (:70,96,126).  The initial content
    // is known up front, so it is sorted once (in parallel, best pair first) and consumed front to back; only the pairs
    // re-queued during clustering live in a real heap.  Taking the larger of "next sorted pair" and "heap top" pops
    // exactly the order one big heap would (the order is total: similarity, then ids), without 14 M cache-missing
    // sift-downs through a 340 MB array.  "Is this pair queued?" needs no 14 M-entry hash set either: an initial pair
    // is queued iff its rank in the sorted run has not been consumed yet (rank_of, found through the sorted, per-vertex
    // indexed key array), and only re-queued pairs go through a small dynamic set.
    const PairOrder before;
    __gnu_parallel::sort(scored.begin(), scored.end(), [&](const Pair &a, const Pair &b) { return before(b, a); });
    std::vector<int> rank_of(npairs);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)npairs; ++r) rank_of[(size_t)scored[r].p1] = (int)r;
    std::vector<int64_t> koff((size_t)numv + 1, 0);  // keys of first endpoint a: [koff[a], koff[a+1])
    for (uint64_t k : keys) ++koff[(size_t)(k >> 32) + 1];
    for (int i = 0; i < numv; ++i) koff[(size_t)i + 1] += koff[i];
    size_t next_sorted = 0;
    auto queued_initial = [&](int x, int y) -> bool {
        const int a = std::min(x, y), b = std::max(x, y);
        const uint64_t key = ((uint64_t)a << 32) | (uint32_t)b;
        const uint64_t *lo = keys.data() + koff[a], *hi = keys.data() + koff[(size_t)a + 1];
        const uint64_t *it = std::lower_bound(lo, hi, key);
        return it != hi && *it == key && (size_t)rank_of[(size_t)(it - keys.data())] >= next_sorted;
    };
    KeySet requeued(1024);
    std::priority_queue<Pair, std::vector<Pair>, PairOrder> heap;  // re-queued root pairs only
    auto put = [&](int p1, int p2) {
        heap.push(Pair{R.jaccard(p1, p2), std::min(p1, p2), std::max(p1, p2), p1, p2});
        requeued.insert(makenum(p1, p2));
    };
    lap("jaccard + pair sort");
    if (trace) fprintf(stderr, "[lsh_reorder] scored pairs: %zu\n", scored.size());
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
    while ((next_sorted < scored.size() || !heap.empty()) && num_cluster > 0) {
        Pair top;
        if (heap.empty() || (next_sorted < scored.size() && !before(scored[next_sorted], heap.top()))) {
            top = scored[next_sorted++];  // the sorted run holds the larger (or the only) candidate
            top.p1 = top.lo, top.p2 = top.hi;
        } else {
            top = heap.top();
            heap.pop();
            requeued.erase(makenum(top.p1, top.p2));
        }
        int p1 = top.p1, p2 = top.p2;
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
            if (p1 != p2 && !requeued.contains(makenum(p1, p2)) && !queued_initial(p1, p2)) put(p1, p2);
        }
    }

    lap("union-find");
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
