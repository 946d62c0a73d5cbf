"""CPU tier: the sampler oracle (oracle.c: orc_sample_subgraph).  fanout <= 0 restates the reference's
sampleVertex (sample.h:131-200) and is checked against a numpy restatement of the same steps (and on the GPU
against the compiled reference, tests/test_gpu_sampler.py); fanout > 0 is the specification of the product's
fixed-fanout sampler (parity unpinned: the reference's sampleVertexSampleNeighbor is not reproducible, see
gnn-computing_b200/csrc/sample_device.cu), checked for the properties it promises."""
import numpy as np
import pytest

from gnnagg import synth


def graph(n, deg, seed, hub=0):
    ptr, idx = synth.small_random_csr(n, deg, seed, empty_frac=0.2, hub=hub)
    return ptr.astype(np.int32), idx.astype(np.int32)


def numpy_sample_vertex(ptr, idx, active, layer_num):
    act = (active != 0).astype(np.int32)
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    for _ in range(layer_num - 1):
        nxt = act.copy()
        nxt[idx[act[rows] != 0]] = 1
        act = nxt
    vs = np.flatnonzero(act).astype(np.int32)
    deg = (ptr[vs + 1] - ptr[vs]).astype(np.int64)
    sp = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    si = np.concatenate([idx[ptr[v]:ptr[v + 1]] for v in vs]).astype(np.int32) if len(vs) else np.empty(0, np.int32)
    return act, vs, sp, si


@pytest.mark.parametrize("layers", [1, 2, 3])
@pytest.mark.parametrize("n,deg,hub", [(1, 2, 0), (60, 3, 0), (500, 6, 300), (40, 0, 0)])
def test_sample_vertex_restatement(orc, n, deg, hub, layers):
    ptr, idx = graph(n, deg, 5, hub)
    rng = np.random.default_rng(n + layers)
    active = (rng.random(n) < 0.1).astype(np.int32) * 7  # any non-zero value marks a seed
    got = orc.sample_subgraph(ptr, idx, active, 0, layers)
    want = numpy_sample_vertex(ptr, idx, active, layers)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_no_seed_and_all_seeds(orc):
    ptr, idx = graph(80, 4, 9)
    act, vs, sp, si = orc.sample_subgraph(ptr, idx, np.zeros(80, np.int32), 0, 3)
    assert act.sum() == 0 and len(vs) == 0 and np.array_equal(sp, [0]) and len(si) == 0
    act, vs, sp, si = orc.sample_subgraph(ptr, idx, np.ones(80, np.int32), 0, 1)
    assert np.array_equal(vs, np.arange(80)) and np.array_equal(sp, ptr) and np.array_equal(si, idx)


@pytest.mark.parametrize("fanout", [1, 4, 16])
def test_fixed_fanout_properties(orc, fanout):
    n = 400
    ptr, idx = graph(n, 10, 3, hub=2000)
    active = np.ones(n, np.int32)
    act, vs, sp, si = orc.sample_subgraph(ptr, idx, active, fanout, 1, seed=123)
    deg = np.diff(ptr)
    assert np.array_equal(vs, np.arange(n))
    assert np.array_equal(np.diff(sp), np.minimum(deg, fanout))
    for v in range(n):
        row, got = idx[ptr[v]:ptr[v + 1]], si[sp[v]:sp[v + 1]]
        if deg[v] <= fanout:
            assert np.array_equal(got, row)
            continue
        pos = [orc.sample_pos(123, v, j, int(deg[v]), fanout) for j in range(fanout)]
        assert all(j * deg[v] // fanout <= p < (j + 1) * deg[v] // fanout for j, p in enumerate(pos))  # one per stratum
        assert len(set(pos)) == fanout and pos == sorted(pos)                                          # distinct, CSR order
        assert np.array_equal(got, row[pos])
    # pure function of (seed, vertex, j): same call, same sample; another seed, another sample
    again = orc.sample_subgraph(ptr, idx, active, fanout, 1, seed=123)
    other = orc.sample_subgraph(ptr, idx, active, fanout, 1, seed=124)
    assert np.array_equal(again[3], si) and not np.array_equal(other[3], si)


def test_fixed_fanout_inclusion_probability(orc):
    """every edge of a long row is kept with probability fanout/deg"""
    deg, fanout, trials = 50, 10, 4000
    hits = np.zeros(deg)
    for s in range(trials):
        for j in range(fanout):
            hits[orc.sample_pos(s, 7, j, deg, fanout)] += 1
    p = hits / trials
    assert abs(p.mean() - fanout / deg) < 1e-12
    assert np.all(np.abs(p - fanout / deg) < 5 * np.sqrt(0.2 * 0.8 / trials))


def test_sampled_expansion_uses_the_same_sample(orc):
    """the hop and the extraction see one sample per vertex (the reference's `expanded`/`chosen` bookkeeping)"""
    n, fanout = 300, 3
    ptr, idx = graph(n, 12, 8)
    active = np.zeros(n, np.int32)
    active[[5, 17, 200]] = 1
    act, vs, sp, si = orc.sample_subgraph(ptr, idx, active, fanout, 2, seed=1)
    # every neighbour kept in the rows of the seeds must itself be active after the hop
    for v in (5, 17, 200):
        r = int(np.searchsorted(vs, v))
        assert vs[r] == v and act[si[sp[r]:sp[r + 1]]].all()
    reached = set(np.flatnonzero(act)) - {5, 17, 200}
    allowed = set()
    for v in (5, 17, 200):
        r = int(np.searchsorted(vs, v))
        allowed |= set(si[sp[r]:sp[r + 1]].tolist())
    assert reached <= allowed
