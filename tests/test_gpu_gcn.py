"""GPU tier: GCN aggregation (gnnagg_gcn_run and friends) through the C ABI vs the CPU oracle.
Tolerance (north_star): fp32 results within 1e-5 relative -- stated per element as
|y - y64| <= 1e-5 * sum_e |val_e * x_e| against the fp64-accumulating oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_gate
from gnnagg import synth
from gpu_util import GRAPHS, dev, make_graph, rand_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("F", [32, 64, 128, 256])
@pytest.mark.parametrize("we", [128, 512])
def test_gcn_unscheduled_parity(gn, orc, cuda, gname, F, we):
    ptr, idx = make_graph(gname, seed=F)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, F, seed=1)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    agg.set_warp_edges(we)  # both item sizes: the automatic choice would only pick 512 above 4M edges
    Y = torch.full((n, F), float("nan"), device=cuda)  # every element must be overwritten
    agg.gcn_run(dev(X), Y)
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    bad, worst = rel_gate(Y.cpu().numpy(), y64, scale, TOL)
    assert bad == 0, (gname, F, worst)
    # reproducible run to run (no float atomics in the un-scheduled path)
    Y2 = torch.empty_like(Y)
    agg.gcn_run(dev(X), Y2)
    assert torch.equal(Y, Y2)


@pytest.mark.parametrize("F", [4, 8, 20, 96, 160, 512, 1024])
def test_gcn_odd_feature_widths(gn, orc, cuda, F):
    ptr, idx = make_graph("medium", seed=3)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, F, seed=2)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    Y = agg.gcn_run(dev(X), torch.empty((n, F), device=cuda))
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(Y.cpu().numpy(), y64, scale, TOL)[0] == 0


def test_gcn_positive_set_plain_relative_error(gn, orc, cuda):
    ptr, idx = make_graph("hub", seed=9)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, 128, seed=4, positive=True)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    Y = agg.gcn_run(dev(X), torch.empty((n, 128), device=cuda)).cpu().numpy()
    y64, _ = orc.spmm_f64(ptr, idx, val, X)
    nz = y64 != 0
    assert np.max(np.abs(Y[nz] - y64[nz]) / np.abs(y64[nz])) <= TOL
    assert np.all(Y[~nz] == 0)


def test_gcn_empty_graph_and_errors(gn, cuda):
    ptr = torch.zeros(8, dtype=torch.int32, device=cuda)
    idx = torch.zeros(0, dtype=torch.int32, device=cuda)
    agg = gn.Aggregator(ptr, idx, torch.zeros(0, device=cuda))
    Y = agg.gcn_run(torch.ones((7, 32), device=cuda), torch.full((7, 32), 5.0, device=cuda))
    assert torch.all(Y == 0)
    with pytest.raises(gn.GnnaggError):
        agg.gcn_run(torch.ones((7, 30), device=cuda), torch.empty((7, 30), device=cuda))  # feat % 4 != 0
    with pytest.raises(gn.GnnaggError):
        agg.gcn_run(torch.ones((7, 32), device=cuda), torch.empty((7, 32), device=cuda), scheduled=True)  # no schedule
    p2, i2 = dev(np.array([0, 1, 2], np.int32)), dev(np.array([1, 0], np.int32))
    agg2 = gn.Aggregator(p2, i2)
    with pytest.raises(gn.GnnaggError):
        agg2.gcn_run(torch.ones((2, 32), device=cuda), torch.empty((2, 32), device=cuda))  # no edge values


@pytest.mark.parametrize("kind,params", [(1, [16]), (1, [32]), (1, [1]), (0, [4]), (2, [4, 32]), (2, [3, 7])])
@pytest.mark.parametrize("F", [32, 128, 256])
def test_gcn_scheduled_parity_and_schedule_upload(gn, orc, cuda, kind, params, F):
    ptr, idx = make_graph("hub", seed=kind * 10 + F)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, F, seed=5)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    agg.set_warp_edges(512 if F == 128 else 0)
    num_target = agg.schedule(kind, params)
    # the uploaded schedule is bit-identical to the reference semantics (oracle)
    if kind == 1:
        ep, ei, et = orc.neighbor_grouping(ptr, idx, params[0])
        ev = val
    else:
        ep, ei, et, ev = orc.locality(ptr, idx, params[0], n, val, neighbor_num=params[1] if kind == 2 else 0)
    sp, si, stt, sv = agg.scheduled_arrays()
    assert num_target == len(et)
    assert np.array_equal(sp, ep) and np.array_equal(si, ei) and np.array_equal(stt, et) and np.array_equal(sv, ev)
    Y = agg.gcn_run(dev(X), torch.full((n, F), float("nan"), device=cuda), scheduled=True)
    y64, scale = orc.spmm_grouped_f64(n, ep, ei, ev, et, X)
    assert rel_gate(Y.cpu().numpy(), y64, scale, TOL)[0] == 0
    # scheduled and un-scheduled agree (same sum, different order)
    y_ref, scale2 = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(Y.cpu().numpy(), y_ref, scale2, TOL)[0] == 0


def test_updateval_after_schedule(gn, orc, cuda):
    """Aggregator_GCN::updateval (aggr_gcn.h:540-544) as used by Figure10/main_a.cu:88,100"""
    ptr, idx = make_graph("medium", seed=1)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, 64, seed=6)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    for kind, params in ((1, [32]), (2, [2, 16])):
        agg.schedule(kind, params)
        val2 = (val * 3 + 1).astype(np.float32)
        agg.set_val(dev(val2))
        Y = agg.gcn_run(dev(X), torch.empty((n, 64), device=cuda), scheduled=True)
        y64, scale = orc.spmm_f64(ptr, idx, val2, X)
        assert rel_gate(Y.cpu().numpy(), y64, scale, TOL)[0] == 0
        agg.set_val(dev(val))


def test_gcn_edgewise_and_edgelist(gn, orc, cuda):
    ptr, idx = make_graph("hub", seed=2)
    n, m = len(ptr) - 1, len(idx)
    for F in (32, 64, 128):
        X, val = rand_inputs(n, m, F, seed=7)
        agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
        Y = agg.gcn_run_edgewise(dev(X), torch.full((n, F), float("nan"), device=cuda))
        y64, scale = orc.spmm_f64(ptr, idx, val, X)
        assert rel_gate(Y.cpu().numpy(), y64, scale, TOL)[0] == 0
    el = agg.csr2edgelist(torch.empty(2 * m, dtype=torch.int32, device=cuda))
    assert np.array_equal(el.cpu().numpy(), orc.csr2edgelist(ptr, idx))


def test_gcn_host_entry_point(gn, orc, cuda):
    ptr, idx = make_graph("medium", seed=4)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, 128, seed=8)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    hX = torch.from_numpy(X).pin_memory()
    hY = torch.empty((n, 128)).pin_memory()
    agg.gcn_run_host(hX, hY)
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(hY.numpy(), y64, scale, TOL)[0] == 0


@pytest.mark.parametrize("F", [32, 128, 256])
@pytest.mark.parametrize("we", [0, 512])
def test_host_entry_points_pipelined_row_chunks(gn, orc, cuda, F, we):
    """above 4096 rows the host-buffer calls aggregate edge-balanced ROW CHUNKS (clipped items) and copy each
    chunk back while the next one computes; hubs, empty chunks' rows and chunk borders inside items included"""
    from gnnagg import synth

    ptr, idx = synth.small_random_csr(9000, 14.0, F, empty_frac=0.3, hub=30000)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, F, seed=3)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    agg.set_warp_edges(we)
    hX = torch.from_numpy(X).pin_memory()
    hY = torch.full((n, F), float("nan")).pin_memory()
    agg.gcn_run_host(hX, hY)
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(hY.numpy(), y64, scale, TOL)[0] == 0
    # identical to the single-launch device path (same items, same order of additions)
    Yd = agg.gcn_run(dev(X), torch.empty((n, F), device=cuda))
    assert np.array_equal(hY.numpy(), Yd.cpu().numpy())
    if F <= 256:
        W = (np.random.default_rng(4).standard_normal((F, 64)) / np.sqrt(F)).astype(np.float32)
        hH = torch.full((n, 64), float("nan")).pin_memory()
        agg.gcn_layer_host(hX, torch.from_numpy(W).pin_memory(), hH)
        _, h64, hs = orc.gcn_layer_f64(ptr, idx, val, X, W)
        assert rel_gate(hH.numpy(), h64, hs, TOL)[0] == 0


@pytest.mark.parametrize("slices", [2, 3, 4, 8])
@pytest.mark.parametrize("F", [32, 128])
def test_host_entry_points_source_slices(gn, orc, cuda, slices, F):
    """large graphs (forced here): X travels in `slices` row blocks and the sub-CSR of the edges whose source lies in a
    block is accumulated as soon as the block is resident; the last slice runs in row chunks.  Hubs, empty rows, a
    source range with no edge at all, new edge values through set_val, repeated calls"""
    from gnnagg import synth

    n = 9001  # not a multiple of any slice count
    ptr, idx = synth.small_random_csr(n, 14.0, F + slices, empty_frac=0.3, hub=30000)
    idx = idx.copy()
    lo, hi = (n // slices) * (slices - 1), n   # leave the last source block almost unused: an (almost) empty slice
    idx[idx >= lo] = idx[idx >= lo] % max(lo, 1)
    ptr, idx = ptr.astype(np.int32), idx.astype(np.int32)
    m = len(idx)
    X, val = rand_inputs(n, m, F, seed=3)
    dval = dev(val)
    agg = gn.Aggregator(dev(ptr), dev(idx), dval)
    agg.set_host_pipeline(slices)
    hX = torch.from_numpy(X).pin_memory()
    hY = torch.full((n, F), float("nan")).pin_memory()
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    for _ in range(2):
        hY.fill_(float("nan"))
        agg.gcn_run_host(hX, hY)
        assert rel_gate(hY.numpy(), y64, scale, TOL)[0] == 0
    first = hY.numpy().copy()
    agg.gcn_run_host(hX, hY)
    assert np.array_equal(first, hY.numpy())  # deterministic
    W = (np.random.default_rng(4).standard_normal((F, 64)) / np.sqrt(F)).astype(np.float32)
    hH = torch.full((n, 64), float("nan")).pin_memory()
    agg.gcn_layer_host(hX, torch.from_numpy(W).pin_memory(), hH)
    _, h64, hs = orc.gcn_layer_f64(ptr, idx, val, X, W)
    assert rel_gate(hH.numpy(), h64, hs, TOL)[0] == 0
    # new edge values are mirrored into slice order
    val2 = (val * 0.25 - 1.0).astype(np.float32)
    dval2 = dev(val2)
    agg.set_val(dval2)
    agg.gcn_run_host(hX, hY)
    y64b, scale_b = orc.spmm_f64(ptr, idx, val2, X)
    assert rel_gate(hY.numpy(), y64b, scale_b, TOL)[0] == 0
    # back to row chunks only: bit-identical to the device path again
    agg.set_host_pipeline(-1)
    agg.gcn_run_host(hX, hY)
    assert np.array_equal(hY.numpy(), agg.gcn_run(dev(X), torch.empty((n, F), device=cuda)).cpu().numpy())


def test_naive_spmm_and_validators(gn, orc, cuda):
    """include/spmm.h: spmm<>, valid(), validReordered()"""
    ptr, idx = make_graph("medium", seed=5)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, 32, seed=9, positive=True)
    Y = torch.full((n, 32), -7.0, device=cuda)
    gn.spmm_naive(dev(ptr), dev(idx), dev(val), dev(X), Y)
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    Yh = Y.cpu().numpy()
    empty = np.diff(ptr) == 0
    assert np.all(Yh[empty] == -7.0)  # untouched, spmm.h:236-237
    assert rel_gate(Yh[~empty], y64[~empty], scale[~empty], TOL)[0] == 0
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    Y2 = agg.gcn_run(dev(X), torch.empty((n, 32), device=cuda))
    ref = dev(y64) + 1.0  # keep away from 0/0
    ans = Y2 + 1.0
    assert gn.validate(ref, ans) == orc.validate2((y64 + 1).astype(np.float32), ans.cpu().numpy()) == 0
    ans[3, 5] += 1.0
    ans[10, 0] *= 1.5
    assert gn.validate(ref, ans) == orc.validate2((y64 + 1).astype(np.float32), ans.cpu().numpy()) == 2
    rows = np.random.default_rng(1).permutation(n).astype(np.int32)
    permuted = torch.empty_like(ans)
    permuted[dev(rows).long()] = ans  # ans row t lives at row rows[t]
    assert gn.validate_reordered(ans, permuted, dev(rows)) == 0
    assert orc.validate_reordered(ans.cpu().numpy(), permuted.cpu().numpy(), rows) == 0
    permuted[rows[2], 1] += 0.5
    assert gn.validate_reordered(ans, permuted, dev(rows)) == 1


@pytest.mark.parametrize("slices", [2, 5, 16])
@pytest.mark.parametrize("F", [32, 256])
def test_locality_slices_match_oracle(gn, orc, cuda, slices, F):
    """gnnagg_set_locality_slices: the un-scheduled aggregation source slice by source slice (first slice complete,
    the others compacted to their non-empty rows and accumulated) -- deterministic, within the parity gate"""
    ptr, idx = synth.small_random_csr(4000, 9.0, 77, empty_frac=0.3, hub=30000)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, F, 5)
    want, scale = orc.spmm_f64(ptr, idx, val, X)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    agg.set_locality_slices(slices)
    Y = agg.gcn_run(dev(X), torch.full((n, F), float("nan"), device=cuda))
    assert rel_gate(Y.cpu().numpy(), want, scale, 1e-5)[0] == 0
    assert torch.equal(Y, agg.gcn_run(dev(X), torch.full((n, F), float("nan"), device=cuda)))   # bit-reproducible
    # accumulate on top of an existing Y, and the fused layer, go through the same slices
    Y2 = agg.gcn_run_acc(dev(X), Y.clone(), accumulate=True)
    assert rel_gate(Y2.cpu().numpy(), 2 * want, 2 * scale, 1e-5)[0] == 0
    if F == 32:
        W = np.random.default_rng(1).standard_normal((F, 64)).astype(np.float32) / 6
        _, h64, hs = orc.gcn_layer_f64(ptr, idx, val, X, W)
        H = agg.gcn_layer(dev(X), dev(W), torch.empty((n, 64), device=cuda))
        assert rel_gate(H.cpu().numpy(), h64, hs, 1e-5)[0] == 0
    agg.set_locality_slices(1)
    assert rel_gate(agg.gcn_run(dev(X), torch.empty((n, F), device=cuda)).cpu().numpy(), want, scale, 1e-5)[0] == 0
