"""GPU tier: fused GAT aggregation, its un-fused pieces, and SDDMM vs the CPU oracle.
GAT tolerance: the kernels use __expf / __fdividef like the reference (--use_fast_math,
CMakeLists.txt:40), the oracle exp() in fp64: 1e-5 of the |.|-sum plus 2e-6 relative for exp."""
import numpy as np
import pytest
import torch

from conftest import rel_gate
from gpu_util import GRAPHS, dev, make_graph, rand_inputs

pytestmark = pytest.mark.gpu
TOL = 1.2e-5


def _att(n, seed):
    return np.random.default_rng(seed).standard_normal((n, 2)).astype(np.float32)


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("F", [32, 64, 128, 256])
@pytest.mark.parametrize("we", [128, 512])
def test_gat_fused_unscheduled(gn, orc, cuda, gname, F, we):
    ptr, idx = make_graph(gname, seed=F + 1)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, F, seed=13)
    att = _att(n, 14)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    agg.set_warp_edges(we)
    Y = agg.gat_run(dev(X), dev(att), torch.full((n, F), float("nan"), device=cuda))
    y64, den, scale = orc.gat_f64(ptr, idx, att, X)  # empty rows -> 0 (documented deviation from NaN)
    bad, worst = rel_gate(Y.cpu().numpy(), y64, scale, TOL)
    assert bad == 0, (gname, F, worst)
    Y2 = agg.gat_run(dev(X), dev(att), torch.empty((n, F), device=cuda))
    assert torch.equal(Y, Y2)  # deterministic


@pytest.mark.parametrize("ng", [16, 32])
@pytest.mark.parametrize("F", [32, 64, 128])
def test_gat_fused_scheduled(gn, orc, cuda, ng, F):
    """aggr_gat_fine + scaleArray (Figure10/main_a.cu:109): result and the un-normalised newval[e]"""
    ptr, idx = make_graph("hub", seed=F + ng)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, F, seed=15)
    att = _att(n, 16)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    agg.set_warp_edges(512 if ng == 16 else 128)
    agg.schedule(1, [ng])
    Y = agg.gat_run(dev(X), dev(att), torch.full((n, F), float("nan"), device=cuda), scheduled=True)
    y64, den, scale = orc.gat_f64(ptr, idx, att, X)
    bad, worst = rel_gate(Y.cpu().numpy(), y64, scale, TOL)
    assert bad == 0, worst
    w = np.empty(m, np.float32)
    gn.check(gn.lib().gnnagg_memcpy_d2h(w.ctypes.data, agg.gat_edge_weights_ptr(), m * 4))
    np.testing.assert_allclose(w, orc.edge_weight_f64(ptr, idx, att), rtol=3e-6)


def test_gat_slope_parameter(gn, orc, cuda):
    ptr, idx = make_graph("medium", seed=8)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, 64, seed=17)
    att = _att(n, 18)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    for slope in (0.2, 0.01, 1.0):
        Y = agg.gat_run(dev(X), dev(att), torch.empty((n, 64), device=cuda), slope=slope)
        y64, _, scale = orc.gat_f64(ptr, idx, att, X, slope=slope)
        assert rel_gate(Y.cpu().numpy(), y64, scale, TOL)[0] == 0


@pytest.mark.parametrize("gname", ["tiny", "short_rows", "hub", "leading_trailing_empty", "exact_items"])
def test_unfused_gat_pieces(gn, orc, cuda, gname):
    """run_att / run_u_add_v / run_add_to_center / run_div_each (aggr_gat.h:395-425)"""
    ptr, idx = make_graph(gname, seed=31)
    n, m = len(ptr) - 1, len(idx)
    att = _att(n, 19)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    datt = dev(att)
    sm = agg.edge_softmax(datt, torch.full((m,), float("nan"), device=cuda)).cpu().numpy()
    np.testing.assert_allclose(sm, orc.edge_softmax_f64(ptr, idx, att), rtol=1e-5, atol=1e-12)
    uv = agg.u_add_v(datt, torch.full((m,), float("nan"), device=cuda))
    assert np.array_equal(uv.cpu().numpy(), orc.u_add_v(ptr, idx, att))  # a single fp32 add: bit-exact
    center = agg.add_to_center(uv, torch.full((n,), float("nan"), device=cuda)).cpu().numpy()
    c64 = orc.add_to_center_f64(ptr, uv.cpu().numpy())
    mag = orc.add_to_center_f64(ptr, np.abs(uv.cpu().numpy()))
    assert np.all(np.abs(center - c64) <= 1e-5 * mag + 1e-30)
    # each_div on values that keep the quotient finite
    pos = torch.rand(m, device=cuda) + 0.5
    csum = agg.add_to_center(pos, torch.empty((n,), device=cuda))
    q = agg.each_div(csum, pos.clone()).cpu().numpy()
    np.testing.assert_allclose(q, orc.each_div(ptr, csum.cpu().numpy(), pos.cpu().numpy()), rtol=2e-7)


def test_unfused_pipeline_equals_fused(gn, orc, cuda):
    """the 'adapter' variant of Figure10/main_a.cu:98-101: attGat -> updateval -> GCN run"""
    ptr, idx = make_graph("hub", seed=41)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, 64, seed=20)
    att = _att(n, 21)
    gat = gn.Aggregator(dev(ptr), dev(idx))
    newval = gat.edge_softmax(dev(att), torch.empty((m,), device=cuda))
    gcn = gn.Aggregator(dev(ptr), dev(idx), newval)
    gcn.schedule(1, [32])
    Y = gcn.gcn_run(dev(X), torch.empty((n, 64), device=cuda), scheduled=True)
    y64, _, scale = orc.gat_f64(ptr, idx, att, X)
    assert rel_gate(Y.cpu().numpy(), y64, scale, 2e-5)[0] == 0


@pytest.mark.parametrize("gname", ["tiny", "short_rows", "hub", "leading_trailing_empty"])
@pytest.mark.parametrize("F", [32, 64, 128, 256, 20])
def test_sddmm(gn, orc, cuda, gname, F):
    ptr, idx = make_graph(gname, seed=51)
    n, m = len(ptr) - 1, len(idx)
    rng = np.random.default_rng(F)
    X1 = rng.standard_normal((n, F)).astype(np.float32)
    X2 = rng.standard_normal((n, F)).astype(np.float32)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    out = agg.sddmm(dev(X1), dev(X2), torch.full((m,), float("nan"), device=cuda)).cpu().numpy()
    v64, scale = orc.sddmm_f64(ptr, idx, X1, X2)
    assert rel_gate(out, v64, scale, 1e-5)[0] == 0
    agg.schedule(1, [16])
    out2 = agg.sddmm(dev(X1), dev(X2), torch.full((m,), float("nan"), device=cuda), scheduled=True).cpu().numpy()
    assert rel_gate(out2, v64, scale, 1e-5)[0] == 0  # NG keeps CSR edge order (graph_schedule.h:123-124)
    agg.schedule(0, [2])
    with pytest.raises(gn.GnnaggError):
        agg.sddmm(dev(X1), dev(X2), torch.empty((m,), device=cuda), scheduled=True)  # aggr_sddmm.h:100
