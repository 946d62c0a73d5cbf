"""cluster2_port.py -- CPU restatement of the reference's locality reorder, script/cluster2.py.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c).  PARITY UNPINNED for the candidate generation:
cluster2.py gets its candidate pairs from datasketch 1.5.1 (MinHash(num_perm=64),
MinHashLSH(threshold=0.2, num_perm=64); script/cluster2.py:2,29-37,85), a dependency that is neither
vendored in /root/reference nor installed here (no network), whose query order depends on
PYTHONHASHSEED, and the heap of cluster2.py breaks similarity ties by insertion history
(:56-67) -- the reference's permutation is not reproducible even by the reference.
What IS restated exactly is everything after the candidates: exact Jaccard (:44-49), the max-heap
greedy union-find with the size cap 64 and the "deleted" freeze (:108-153), the re-queueing of root
pairs scored on the roots' own lists (:144-153), and the output order (:156-171).  That control flow
is pinned by the 8-vertex run of the real cluster2.py recorded in SURVEY.md section 4
(tests/test_reorder.py).  The candidate generation is re-specified deterministically:

  hash of a neighbour id     h(u)   = low 32 bits of splitmix64(u)            (datasketch: sha1 of str(u))
  permutation k              pi_k(x) = ((a_k*x + b_k) mod (2^61-1)) & 0xffffffff, a_k,b_k from splitmix64(seed,k)
  signature                  sig_i[k] = min_u pi_k(h(u)), 0xffffffff for an empty row
  LSH                        b bands of r rows (datasketch's optimum for threshold 0.2 / 64 perms: b=28, r=2);
                             vertices sharing a band bucket are candidates
  bucket window              inside a bucket (members sorted by id) only members at most WINDOW positions
                             apart are paired (identical rows of a hub's leaves form buckets of 10^4+)
  candidate order            ascending id; heap key (-similarity, min id, max id)
"""
import heapq

P61 = (1 << 61) - 1
M64 = (1 << 64) - 1
WINDOW = 32
MAXH = 0xFFFFFFFF


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def perm_params(num_perm, seed):
    a = [1 + splitmix64((seed << 32) + 2 * k) % (P61 - 1) for k in range(num_perm)]
    b = [splitmix64((seed << 32) + 2 * k + 1) % P61 for k in range(num_perm)]
    return a, b


def signatures(ptr, idx, num_perm, seed):
    a, b = perm_params(num_perm, seed)
    sigs = []
    for i in range(len(ptr) - 1):
        sig = [MAXH] * num_perm
        for e in range(ptr[i], ptr[i + 1]):
            h = splitmix64(int(idx[e])) & MAXH
            for k in range(num_perm):
                v = ((a[k] * h + b[k]) % P61) & MAXH
                if v < sig[k]:
                    sig[k] = v
        sigs.append(sig)
    return sigs


def jaccard(l1, l2):
    """script/cluster2.py:44-49"""
    if len(l1) == 0 or len(l2) == 0:
        return 0.0
    s1, s2 = set(l1), set(l2)
    return float(len(s1 & s2)) / len(s1 | s2)


def candidates_lsh(ptr, idx, num_perm, bands, rows_per_band, seed):
    numv = len(ptr) - 1
    sigs = signatures(ptr, idx, num_perm, seed)
    cand = [set() for _ in range(numv)]
    for j in range(bands):
        buckets = {}
        for i in range(numv):
            buckets.setdefault(tuple(sigs[i][j * rows_per_band:(j + 1) * rows_per_band]), []).append(i)
        for members in buckets.values():  # members ascending by construction
            for p, i in enumerate(members):
                for q in range(max(0, p - WINDOW), min(len(members), p + WINDOW + 1)):
                    if q != p:
                        cand[i].add(members[q])
    return [sorted(c) for c in cand]


def candidates_exhaustive(ptr, idx):
    """every pair of vertices that share a neighbour (what a perfect LSH would return); small graphs only"""
    numv = len(ptr) - 1
    owners = {}
    for i in range(numv):
        for e in range(ptr[i], ptr[i + 1]):
            owners.setdefault(int(idx[e]), set()).add(i)
    cand = [set() for _ in range(numv)]
    for members in owners.values():
        for i in members:
            cand[i] |= members
    return [sorted(c - {i}) for i, c in enumerate(cand)]


def cluster(ptr, idx, num_perm=64, bands=28, rows_per_band=2, cap=64, seed=1):
    """returns rows[numv]: entry k = old id placed at new position k.  bands < 0: exhaustive candidates"""
    numv = len(ptr) - 1
    lists = [[int(x) for x in idx[ptr[i]:ptr[i + 1]]] for i in range(numv)]
    cand = candidates_exhaustive(ptr, idx) if bands < 0 else candidates_lsh(ptr, idx, num_perm, bands, rows_per_band, seed)

    def makenum(a, b):  # cluster2.py:73-78
        return (a * numv + b) if a <= b else (b * numv + a)

    heap, sset = [], set()

    def put(p1, p2):
        heapq.heappush(heap, (-jaccard(lists[p1], lists[p2]), min(p1, p2), max(p1, p2), p1, p2))
        sset.add(makenum(p1, p2))

    for i in range(numv):  # cluster2.py:80-96
        if ptr[i] == ptr[i + 1]:
            continue
        for c in cand[i]:
            if c == i or makenum(i, c) in sset:
                continue
            put(i, c)

    cluster_id = list(range(numv))
    cluster_sz = [1] * numv
    deleted = [0] * numv

    def root(i):  # cluster2.py:112-116
        while i != cluster_id[i]:
            cluster_id[i] = cluster_id[cluster_id[i]]
            i = cluster_id[i]
        return i

    num_cluster = numv
    while heap and num_cluster > 0:  # cluster2.py:121-153
        _, _, _, p1, p2 = heapq.heappop(heap)
        sset.remove(makenum(p1, p2))
        if p1 == cluster_id[p1] and p2 == cluster_id[p2]:
            if deleted[p1] or deleted[p2]:
                continue
            if cluster_sz[p1] < cluster_sz[p2]:
                cluster_id[p1] = p2
                num_cluster -= 1
                cluster_sz[p2] += cluster_sz[p1]
                if cluster_sz[p2] >= cap:
                    deleted[p2] = 1
                    num_cluster -= 1
            else:
                cluster_id[p2] = p1
                num_cluster -= 1
                cluster_sz[p1] += cluster_sz[p2]
                if cluster_sz[p1] >= cap:
                    deleted[p1] = 1
                    num_cluster -= 1
        else:
            p1, p2 = root(p1), root(p2)
            if deleted[p1] or deleted[p2]:
                continue
            if p1 != p2 and makenum(p1, p2) not in sset:
                put(p1, p2)

    clusters = {}  # cluster2.py:156-164 (dict keeps first-insertion order)
    for i in range(numv):
        clusters.setdefault(root(i), []).append(i)
    return [v for members in clusters.values() for v in members]
