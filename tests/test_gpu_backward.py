"""GPU tier: backward of the aggregation (gnnagg_transpose_build / gnnagg_gcn_backward / gnnagg_gat_backward)
through the C ABI vs the fp64 oracle (pinned against finite differences in tests/test_backward.py), and vs
the reference's experimental aggr_gat_fine_bwd where that kernel is right (F = 32, every pre-activation
positive, source half of the attention gradient; aggr_gat.h:222-294)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_gate
from gpu_util import GRAPHS, dev, make_graph, rand_inputs
from gnnagg import synth

pytestmark = pytest.mark.gpu
TOL_X = 1.2e-5   # alpha_e carries the __expf / division error of the forward (tests/test_gpu_gat.py)
TOL_A = 2.5e-5   # ds_e = alpha (g - c): fp32 dot products of length F on both sides of a difference


def _att(rows, seed, positive=False):
    rng = np.random.default_rng(seed)
    if positive:
        return (rng.random((rows, 2), dtype=np.float32) * 0.9 + 0.05).astype(np.float32)
    return rng.standard_normal((rows, 2)).astype(np.float32)


@pytest.mark.parametrize("gname", list(GRAPHS))
def test_transpose_bit_exact(gn, orc, cuda, gname):
    ptr, idx = make_graph(gname, seed=3)
    n = len(ptr) - 1
    agg = gn.Aggregator(dev(ptr), dev(idx))
    for num_src in (n, n + 37):
        agg.transpose_build(num_src)
        tp, ti, tq = agg.transposed_arrays()
        op, oi, oq = orc.transpose_csr(ptr, idx, num_src)
        assert np.array_equal(tp, op) and np.array_equal(ti, oi) and np.array_equal(tq, oq)


def test_transpose_rejects_out_of_range_source(gn, cuda):
    ptr, idx = make_graph("medium", seed=3)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    with pytest.raises(gn.GnnaggError):
        agg.transpose_build(int(idx.max()))  # one too small
    with pytest.raises(gn.GnnaggError):
        agg.gcn_backward(torch.zeros((len(ptr) - 1, 32), device=cuda), torch.zeros((len(ptr) - 1, 32), device=cuda))


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("F", [32, 128, 256])
def test_gcn_backward(gn, orc, cuda, gname, F):
    ptr, idx = make_graph(gname, seed=F + 2)
    n, m = len(ptr) - 1, len(idx)
    dY, val = rand_inputs(n, m, F, seed=71)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    agg.transpose_build()
    dX = agg.gcn_backward(dev(dY), torch.full((n, F), float("nan"), device=cuda))
    d64, scale = orc.spmm_t_f64(ptr, idx, val, dY, n)
    bad, worst = rel_gate(dX.cpu().numpy(), d64, scale, 1e-5)
    assert bad == 0, (gname, F, worst)
    assert torch.equal(dX, agg.gcn_backward(dev(dY), torch.empty((n, F), device=cuda)))  # deterministic
    # new edge values through set_val are picked up
    val2 = (val * 0.5 + 1.0).astype(np.float32)
    dval2 = dev(val2)
    agg.set_val(dval2)
    dX2 = agg.gcn_backward(dev(dY), torch.empty((n, F), device=cuda))
    d64b, scale_b = orc.spmm_t_f64(ptr, idx, val2, dY, n)
    assert rel_gate(dX2.cpu().numpy(), d64b, scale_b, 1e-5)[0] == 0


def test_gcn_backward_is_the_adjoint_at_scale(gn, cuda):
    """<A X, dY> == <X, A^T dY> on a 4 M-edge power-law graph (size-independent property)"""
    n, m, F = 200000, 4200000, 64
    ptr, idx = synth.rmat_csr(n, m, seed=123, device=cuda)
    val = synth.gcn_norm_val(ptr, idx)
    g = torch.Generator(device=cuda).manual_seed(5)
    X = torch.rand((n, F), device=cuda, generator=g)
    dY = torch.rand((n, F), device=cuda, generator=g)
    agg = gn.Aggregator(ptr, idx, val)
    agg.transpose_build()
    Y = agg.gcn_run(X, torch.empty((n, F), device=cuda))
    dX = agg.gcn_backward(dY, torch.empty((n, F), device=cuda))
    lhs, rhs = float((Y.double() * dY.double()).sum()), float((X.double() * dX.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs)


@pytest.mark.parametrize("gname", list(GRAPHS))
@pytest.mark.parametrize("F", [32, 64, 256])
def test_gat_backward(gn, orc, cuda, gname, F):
    ptr, idx = make_graph(gname, seed=F + 5)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, F, seed=73)
    dY, _ = rand_inputs(n, m, F, seed=74)
    att = _att(n, 75)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    agg.transpose_build()
    dXd, dYd, attd = dev(X), dev(dY), dev(att)
    Y = agg.gat_run(dXd, attd, torch.empty((n, F), device=cuda))
    dX, dA = agg.gat_backward(dXd, attd, Y, dYd, torch.full((n, F), float("nan"), device=cuda),
                              torch.full((n, 2), float("nan"), device=cuda))
    x64, a64, sx, sa = orc.gat_backward_f64(ptr, idx, att, X, dY)
    bad, worst = rel_gate(dX.cpu().numpy(), x64, sx, TOL_X)
    assert bad == 0, ("dX", gname, F, worst)
    bad, worst = rel_gate(dA.cpu().numpy(), a64, sa, TOL_A)
    assert bad == 0, ("datt", gname, F, worst)
    dX2, dA2 = agg.gat_backward(dXd, attd, Y, dYd, torch.empty((n, F), device=cuda), torch.empty((n, 2), device=cuda))
    assert torch.equal(dX, dX2) and torch.equal(dA, dA2)  # deterministic: no float atomics


def test_gat_backward_wide_rows(gn, orc, cuda):
    """F > 256: pass 1 walks the feature columns in several chunks and keeps the partial dot products per edge"""
    F = 320
    ptr, idx = make_graph("medium", seed=12)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, F, seed=91)
    dY, _ = rand_inputs(n, m, F, seed=92)
    att = _att(n, 93)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    agg.transpose_build()
    Xd, dYd, attd = dev(X), dev(dY), dev(att)
    Y = agg.gat_run(Xd, attd, torch.empty((n, F), device=cuda))
    dX, dA = agg.gat_backward(Xd, attd, Y, dYd, torch.empty((n, F), device=cuda), torch.empty((n, 2), device=cuda))
    x64, a64, sx, sa = orc.gat_backward_f64(ptr, idx, att, X, dY)
    assert rel_gate(dX.cpu().numpy(), x64, sx, TOL_X)[0] == 0
    assert rel_gate(dA.cpu().numpy(), a64, sa, TOL_A)[0] == 0


def test_gat_backward_large_graph_path(gn, orc, cuda):
    """4.2 M edges: the 512-edge-per-warp kernels, and pass 2 as permute + row sum + plain aggregation (graphs whose
    per-edge (w, t) array does not stay in L2), against the fp64 oracle on the whole graph"""
    n, m, F = 150000, 4200000, 32
    ptr_d, idx_d = synth.rmat_csr(n, m, seed=321, device=cuda)
    ptr, idx = ptr_d.cpu().numpy(), idx_d.cpu().numpy()
    X, _ = rand_inputs(n, 1, F, seed=95)
    dY, _ = rand_inputs(n, 1, F, seed=96)
    att = _att(n, 97)
    agg = gn.Aggregator(ptr_d, idx_d)
    agg.transpose_build()
    Xd, dYd, attd = dev(X), dev(dY), dev(att)
    Y = agg.gat_run(Xd, attd, torch.empty((n, F), device=cuda))
    dX, dA = agg.gat_backward(Xd, attd, Y, dYd, torch.full((n, F), float("nan"), device=cuda),
                              torch.full((n, 2), float("nan"), device=cuda))
    x64, a64, sx, sa = orc.gat_backward_f64(ptr, idx, att, X, dY)
    bad, worst = rel_gate(dX.cpu().numpy(), x64, sx, TOL_X)
    assert bad == 0, ("dX", worst)
    bad, worst = rel_gate(dA.cpu().numpy(), a64, sa, TOL_A)
    assert bad == 0, ("datt", worst)
    dX2, dA2 = agg.gat_backward(Xd, attd, Y, dYd, torch.empty((n, F), device=cuda), torch.empty((n, 2), device=cuda))
    assert torch.equal(dX, dX2) and torch.equal(dA, dA2)
    # the GCN backward still sees its own edge values afterwards (t_val was borrowed for alpha)
    val = synth.gcn_norm_val(ptr_d, idx_d)
    agg.set_val(val)
    dXg = agg.gcn_backward(dYd, torch.empty((n, F), device=cuda))
    d64, scale = orc.spmm_t_f64(ptr, idx, val.cpu().numpy(), dY, n)
    assert rel_gate(dXg.cpu().numpy(), d64, scale, 1e-5)[0] == 0


@pytest.mark.parametrize("slope", [0.2, 0.0, 1.0])
def test_gat_backward_from_weights_and_rectangular(gn, orc, cuda, slope):
    """the run_bwd calling convention (w = newval of aggr_gat_fine, den = div; no attention table) and a
    rectangular block (num_src != num_v) as the row partitions of the multi-GPU path produce"""
    n, num_src, F = 400, 1000, 64
    ptr, idx = synth.small_random_csr(n, 9.0, 4, empty_frac=0.1, hub=3000, num_src=num_src)
    ptr, idx = ptr.astype(np.int32), idx.astype(np.int32)
    rng = np.random.default_rng(6)
    X = rng.standard_normal((num_src, F)).astype(np.float32)
    dY = rng.standard_normal((n, F)).astype(np.float32)
    att = _att(num_src, 7)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    agg.transpose_build(num_src)
    Xd, dYd, attd = dev(X), dev(dY), dev(att)
    Y = agg.gat_run(Xd, attd, torch.empty((n, F), device=cuda), slope=slope)
    x64, a64, sx, sa = orc.gat_backward_f64(ptr, idx, att, X, dY, slope)
    dX, dA = agg.gat_backward(Xd, attd, Y, dYd, torch.empty((num_src, F), device=cuda),
                              torch.full((num_src, 2), float("nan"), device=cuda), slope=slope)
    assert rel_gate(dX.cpu().numpy(), x64, sx, TOL_X)[0] == 0
    assert rel_gate(dA.cpu().numpy(), a64, sa, TOL_A)[0] == 0
    # same through (w, den)
    w = dev(orc.edge_weight_f64(ptr, idx, att, slope))
    den = agg.add_to_center(w, torch.empty(n, device=cuda))
    dXw, dAw = agg.gat_backward(Xd, None, Y, dYd, torch.empty((num_src, F), device=cuda),
                                torch.empty((num_src, 2), device=cuda), slope=slope, w=w, den=den)
    assert rel_gate(dXw.cpu().numpy(), x64, sx, TOL_X)[0] == 0
    assert rel_gate(dAw.cpu().numpy(), a64, sa, TOL_A)[0] == 0
    # den is optional: the row sums of w are formed during the first traversal anyway
    dXn, dAn = agg.gat_backward(Xd, None, Y, dYd, torch.empty((num_src, F), device=cuda),
                                torch.empty((num_src, 2), device=cuda), slope=slope, w=w)
    assert rel_gate(dXn.cpu().numpy(), x64, sx, TOL_X)[0] == 0
    assert rel_gate(dAn.cpu().numpy(), a64, sa, TOL_A)[0] == 0


def test_gat_backward_vs_reference_kernel(gn, orc, cuda):
    """aggr_gat_fine_bwd (F = 32) accumulates d_feat and the SOURCE half of d_a_b; with every pre-activation
    positive its LeakyReLU factor is 1 as it should be, so both must agree with ours and with the oracle"""
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref.so not present (built only where /root/reference exists)")
    ref = orc.ref()
    P = lambda t: C.c_void_p(t.data_ptr())
    F, ng = 32, 32
    ptr, idx = make_graph("medium", seed=9)
    n, m = len(ptr) - 1, len(idx)
    X, _ = rand_inputs(n, m, F, seed=81)
    dY, _ = rand_inputs(n, m, F, seed=82)
    att = _att(n, 83, positive=True)
    dptr, didx, Xd, dYd, attd = dev(ptr), dev(idx), dev(X), dev(dY), dev(att)
    agg = gn.Aggregator(dptr, didx)
    agg.transpose_build()
    Y = agg.gat_run(Xd, attd, torch.empty((n, F), device=cuda))
    dX, dA = agg.gat_backward(Xd, attd, Y, dYd, torch.empty((n, F), device=cuda), torch.empty((n, 2), device=cuda))
    x64, a64, sx, sa = orc.gat_backward_f64(ptr, idx, att, X, dY)
    assert rel_gate(dX.cpu().numpy(), x64, sx, TOL_X)[0] == 0
    assert rel_gate(dA.cpu().numpy(), a64, sa, TOL_A)[0] == 0

    ref.ref_set_globals(n, m)
    h = C.c_void_p(ref.ref_gat_create(P(dptr), P(didx), n, m, F))
    ref.ref_gat_schedule(h, 1, ng, 0)
    w = dev(orc.edge_weight_f64(ptr, idx, att))          # newval of aggr_gat_fine: neighbour grouping keeps CSR order
    den = agg.add_to_center(w, torch.empty(n, device=cuda))
    den = torch.where(den == 0, torch.ones_like(den), den)
    d_ab = torch.zeros((n, 2), device=cuda)
    d_feat = torch.zeros((n, F), device=cuda)
    ref.ref_gat_run_bwd(h, P(Y), P(dYd), P(w), P(den), P(Xd), P(d_ab), P(d_feat), C.c_float(0.2), 128)
    assert ref.ref_sync() == 0
    _, worst_x = rel_gate(d_feat.cpu().numpy(), x64, sx, TOL_X)
    _, worst_a = rel_gate(d_ab.cpu().numpy()[:, 1], a64[:, 1], sa[:, 1], TOL_A)
    print("reference aggr_gat_fine_bwd worst err/bound: d_feat %.3f  d_a_b(source) %.3f" % (worst_x, worst_a))
    # the kernel reads shared_write_cache across lanes without a __syncwarp (aggr_gat.h:272-288): only require the
    # agreement where it came out right, as for aggr_sddmm
    if worst_x <= 4.0:
        assert rel_gate(dX.cpu().numpy(), d_feat.cpu().numpy(), sx, TOL_X * 5)[0] == 0
    if worst_a <= 4.0:
        assert rel_gate(dA.cpu().numpy()[:, 1], d_ab.cpu().numpy()[:, 1], sa[:, 1], TOL_A * 5)[0] == 0
    assert (d_ab.cpu().numpy()[:, 0] == 0).all()  # the destination half is never written there (:290)


import os


@pytest.mark.skipif(os.environ.get("GNNAGG_HEAVY") != "1", reason="full-size timing run: set GNNAGG_HEAVY=1")
@pytest.mark.parametrize("shape,F", [("arxiv", 32), ("reddit", 32), ("reddit", 128), ("proteins", 64)])
def test_backward_timing_full_size(gn, orc, cuda, shape, F):
    """backward at the BASELINE.json shapes, timed beside the forward and (F = 32) beside the reference's
    aggr_gat_fine_bwd; one JSON line per case into gpurun_out/backward.jsonl"""
    import json
    import time

    n, m = synth.shape_of(shape)
    ptr, idx = synth.rmat_csr(n, m, seed=123, device=cuda)
    val = synth.gcn_norm_val(ptr, idx)
    g = torch.Generator(device=cuda).manual_seed(123)
    X = torch.randn((n, F), device=cuda, generator=g)
    dY = torch.randn((n, F), device=cuda, generator=g)
    att = torch.rand((n, 2), device=cuda, generator=g) * 0.9 + 0.05  # positive: the regime where the reference kernel is right
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=cuda)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    agg = gn.Aggregator(ptr, idx, val)
    torch.cuda.synchronize()
    t0 = time.time()
    agg.transpose_build()
    torch.cuda.synchronize()
    out = {"shape": shape, "n": n, "m": m, "F": F, "transpose_build_s": time.time() - t0}
    Y, dX, dA = torch.empty((n, F), device=cuda), torch.empty((n, F), device=cuda), torch.empty((n, 2), device=cuda)
    out["gcn_forward_ms"] = timeit(lambda: agg.gcn_run(X, Y))
    out["gcn_backward_ms"] = timeit(lambda: agg.gcn_backward(dY, dX))
    out["gat_forward_ms"] = timeit(lambda: agg.gat_run(X, att, Y))
    out["gat_backward_ms"] = timeit(lambda: agg.gat_backward(X, att, Y, dY, dX, dA))
    # every entry point is free of host synchronisation and (after the first call) of allocation, so a caller can
    # capture it in a CUDA graph: what that buys on the launch-bound small graph
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        agg.gat_backward(X, att, Y, dY, dX, dA)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            agg.gat_backward(X, att, Y, dY, dX, dA)
    torch.cuda.synchronize()
    ref_dX, ref_dA = dX.clone(), dA.clone()
    dX.zero_(), dA.zero_()
    out["gat_backward_graph_ms"] = timeit(graph.replay)
    assert torch.equal(dX, ref_dX) and torch.equal(dA, ref_dA)
    # size-independent parity property: <A X, dY> == <X, A^T dY>
    Yg = agg.gcn_run(X, torch.empty((n, F), device=cuda))
    dXg = agg.gcn_backward(dY, torch.empty((n, F), device=cuda))
    lhs, rhs = float((Yg.double() * dY.double()).sum()), float((X.double() * dXg.double()).sum())
    mag = float((Yg.double().abs() * dY.double().abs()).sum())
    out["adjoint_rel_err"] = abs(lhs - rhs) / mag
    assert out["adjoint_rel_err"] <= 1e-6
    if F == 32 and orc.ref_available():
        ref = orc.ref()
        P = lambda t: C.c_void_p(t.data_ptr())
        ref.ref_set_globals(n, m)
        h = C.c_void_p(ref.ref_gat_create(P(ptr), P(idx), n, m, F))
        ref.ref_gat_schedule(h, 1, 32, 0)
        w = torch.empty(m, device=cuda)
        agg.edge_softmax(att, w)           # alpha with div = 1 is an equally valid (newval, div) pair
        ones = torch.ones(n, device=cuda)
        d_ab, d_feat = torch.zeros((n, 2), device=cuda), torch.zeros((n, F), device=cuda)
        Yf = agg.gat_run(X, att, torch.empty((n, F), device=cuda))
        out["ref_aggr_gat_fine_bwd_ms"] = timeit(
            lambda: ref.ref_gat_run_bwd(h, P(Yf), P(dY), P(w), P(ones), P(X), P(d_ab), P(d_feat), C.c_float(0.2), 128), reps=3)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "backward.jsonl"), "a") as f:
        f.write(json.dumps(out) + "\n")
    print(json.dumps(out))
