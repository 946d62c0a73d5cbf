"""CPU tier: locality reorder (gnnagg_lsh_reorder, re-statement of script/cluster2.py).
Parity status: candidate generation is PARITY UNPINNED (datasketch absent and non-deterministic, see
oracle/cluster2_port.py); the clustering control flow is pinned by the recorded 8-vertex run of the
real cluster2.py and the product is held bit-equal to the executable spec."""
import numpy as np
import pytest

from gnnagg import synth
from oracle import cluster2_port as cp


def _csr(rows):
    ptr = np.zeros(len(rows) + 1, np.int32)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    return ptr, np.array(sum(rows, []), np.int32)


def test_recorded_cluster2_run(gn):
    """cluster2.py on this graph wrote '0 1 2 4 5 6 7 3 '"""
    ptr, idx = _csr([[1, 2, 4], [0, 2], [0, 1], [], [5, 6], [4, 6, 7], [4, 5], [5]])
    assert cp.cluster(ptr, idx, bands=-1) == [0, 1, 2, 4, 5, 6, 7, 3]
    assert gn.lsh_reorder(ptr, idx, bands=-1).tolist() == [0, 1, 2, 4, 5, 6, 7, 3]
    assert gn.lsh_reorder(ptr, idx).tolist() == cp.cluster(ptr, idx)


@pytest.mark.parametrize("seed", range(5))
def test_product_equals_spec_random(gn, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(20, 160))
    ptr, idx = synth.small_random_csr(n, float(rng.uniform(1, 6)), seed, empty_frac=0.2, num_src=max(4, n // 4))
    for kw in (dict(bands=-1), dict(), dict(num_perm=16, bands=4, rows_per_band=3, cluster_cap=5, seed=9)):
        got = gn.lsh_reorder(ptr, idx, **kw).tolist()
        want = cp.cluster(ptr, idx, **{("cap" if k == "cluster_cap" else k): v for k, v in kw.items()})
        assert got == want, kw
        assert sorted(got) == list(range(n))  # a permutation


def test_planted_clusters_become_contiguous_and_cap_holds(gn):
    """vertices with identical neighbour lists must end up adjacent; no cluster absorbs after reaching the cap"""
    rng = np.random.default_rng(3)
    groups, per = 12, 20
    protos = [list(range(g * 10, g * 10 + 8)) for g in range(groups)]  # disjoint: no similarity across groups
    owner = rng.permutation(np.repeat(np.arange(groups), per))
    ptr, idx = _csr([protos[g] for g in owner])
    rows = gn.lsh_reorder(ptr, idx)
    assert sorted(rows.tolist()) == list(range(groups * per))
    g_in_order = owner[rows]
    changes = int((np.diff(g_in_order) != 0).sum())
    assert changes == groups - 1  # each planted group is one contiguous run
    # cap: 100 identical rows with cap 8 -> frozen clusters stop growing (cluster2.py:134-143)
    ptr, idx = _csr([[1, 2, 3]] * 100)
    rows = gn.lsh_reorder(ptr, idx, cluster_cap=8)
    assert rows.tolist() == cp.cluster(ptr, idx, cap=8)


def test_reorder_roundtrip_through_loader(gn, orc, tmp_path):
    """the permutation file produced here is what load_graph consumes (src/data.cu:96-133)"""
    ptr, idx = synth.small_random_csr(300, 4.0, 11, num_src=300)
    rows = gn.lsh_reorder(ptr, idx)
    d = str(tmp_path) + "/"
    gn.write_graph("g", ptr, idx, d)
    gn.write_reorder(d + "g.reorder_thres_0.2", rows)
    p2, i2, r2, rev2 = gn.load_graph("g", d, "_thres_0.2")
    assert np.array_equal(r2, rows)
    ep, ei = orc.reorder_csr(ptr, idx, rows, rev2)
    assert np.array_equal(p2, ep) and np.array_equal(i2, ei)
    # relabelled graph is isomorphic: degree multiset and edge count preserved
    assert sorted(np.diff(p2).tolist()) == sorted(np.diff(ptr).tolist())
