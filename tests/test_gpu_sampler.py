"""GPU tier: sub-graph samplers through the C ABI (gnnagg_sample_subgraph) -- bit-exact against the CPU oracle in
both modes, bit-exact against the reference's own sampleVertex (compiled into oracle/_ref/libref.so), and the
aggregation over a sampled CSRSubGraph against the same rows of the full-graph aggregation."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import GRAPHS, dev, make_graph, rand_inputs
from gnnagg import synth

pytestmark = pytest.mark.gpu


def seeds(n, frac, seed):
    return (np.random.default_rng(seed).random(n) < frac).astype(np.int32)


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("fanout", [0, 1, 5, 16])
@pytest.mark.parametrize("layers", [1, 2, 3])
def test_sampler_matches_oracle(gn, orc, cuda, gname, fanout, layers):
    ptr, idx = make_graph(gname, seed=layers)
    n = len(ptr) - 1
    active = seeds(n, 0.15, 3) * 5
    agg = gn.Aggregator(dev(ptr), dev(idx))
    d_active = dev(active)
    vs, sp, si = agg.sample_subgraph(d_active, fanout, layers, seed=77)
    o_act, o_vs, o_sp, o_si = orc.sample_subgraph(ptr, idx, active, fanout, layers, seed=77)
    assert np.array_equal(d_active.cpu().numpy(), o_act)
    assert np.array_equal(vs.cpu().numpy(), o_vs)
    assert np.array_equal(sp.cpu().numpy(), o_sp)
    assert np.array_equal(si.cpu().numpy(), o_si)


@pytest.mark.parametrize("layers", [1, 2, 3])
def test_sample_vertex_matches_reference(gn, orc, cuda, layers):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref.so not present (built only where /root/reference exists)")
    ref = orc.ref()
    P = lambda t: C.c_void_p(t.data_ptr())
    ptr, idx = make_graph("hub", seed=4)
    n, m = len(ptr) - 1, len(idx)
    active = seeds(n, 0.05, 11)
    dptr, didx = dev(ptr), dev(idx)
    ref.ref_set_globals(n, m)
    r_active = dev(active)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    vs_p, sp_p, si_p, ne = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
    nv = ref.ref_sample_vertex(P(r_active), P(dptr), P(didx), layers, C.byref(vs_p), C.byref(sp_p), C.byref(si_p), C.byref(ne))
    assert ref.ref_sync() == 0

    def back(p, count):
        out = np.empty(count, np.int32)
        if count:
            gn.check(gn.lib().gnnagg_memcpy_d2h(out.ctypes.data, p, count * 4))
        return out

    r_vs, r_sp, r_si = back(vs_p, nv), back(sp_p, nv + 1), back(si_p, ne.value)
    agg = gn.Aggregator(dptr, didx)
    o_active = dev(active)
    vs, sp, si = agg.sample_subgraph(o_active, 0, layers)
    assert np.array_equal(o_active.cpu().numpy(), r_active.cpu().numpy())
    assert np.array_equal(vs.cpu().numpy(), r_vs)
    # The reference scans `sub_degree` without ever zeroing the entries of inactive vertices (sample.h:166 allocates,
    # getSubDegree :38-45 writes active entries only), so its row pointers are right only when cudaMalloc happens to
    # return zeroed memory.  Compare pointers and edges when its own output is self-consistent; say so otherwise.
    deg = (ptr[r_vs + 1] - ptr[r_vs]).astype(np.int64)
    consistent = nv > 0 and ne.value == int(deg.sum()) and np.array_equal(np.diff(r_sp), deg)
    print("reference sampleVertex self-consistent:", consistent)
    if consistent:
        assert np.array_equal(sp.cpu().numpy(), r_sp)
        assert np.array_equal(si.cpu().numpy(), r_si)
    else:
        o = orc.sample_subgraph(ptr, idx, active, 0, layers)
        assert np.array_equal(sp.cpu().numpy(), o[2]) and np.array_equal(si.cpu().numpy(), o[3])


@pytest.mark.parametrize("fanout", [0, 8])
def test_aggregation_over_sampled_subgraph(gn, orc, cuda, fanout):
    """CSRSubGraph -> Aggregator (Figure8-style use): row r of the result is the aggregation of vertexset[r]'s kept
    neighbours, gathered from the FULL feature matrix (global source ids)"""
    n, m, F = 30000, 600000, 64
    ptr, idx = synth.rmat_csr(n, m, seed=123, device=cuda)
    g = torch.Generator(device=cuda).manual_seed(1)
    X = torch.randn((n, F), device=cuda, generator=g)
    full = gn.Aggregator(ptr, idx)
    active = (torch.rand(n, device=cuda, generator=g) < 0.02).to(torch.int32)
    vs, sp, si = full.sample_subgraph(active, fanout, 2, seed=9)
    val = torch.ones(si.numel(), device=cuda)
    sub = gn.Aggregator(sp, si, val)
    Y = sub.gcn_run(X, torch.empty((vs.numel(), F), device=cuda))
    y64, scale = orc.spmm_f64(sp.cpu().numpy(), si.cpu().numpy(), val.cpu().numpy(), X.cpu().numpy())
    from conftest import rel_gate

    assert rel_gate(Y.cpu().numpy(), y64, scale, 1e-5)[0] == 0
    if fanout == 0:  # complete rows: identical to the same rows of the full-graph aggregation
        Yf = gn.Aggregator(ptr, idx, torch.ones(m, device=cuda)).gcn_run(X, torch.empty((n, F), device=cuda))
        assert rel_gate(Y.cpu().numpy(), Yf[vs.long()].cpu().numpy(), scale, 2e-5)[0] == 0


def test_sampler_rejects_bad_arguments(gn, cuda):
    ptr, idx = make_graph("medium", seed=1)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    with pytest.raises(gn.GnnaggError):
        agg.sample_subgraph(dev(np.ones(len(ptr) - 1, np.int32)), 0, 0)
