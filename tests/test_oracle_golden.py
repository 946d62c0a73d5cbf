"""CPU tier: the oracle (oracle/oracle.c) against the golden vectors produced by the reference's
own host functions (tests/golden/make_golden.py) and, when oracle/_ref/libref.so is present,
against the compiled reference itself on random graphs."""
import os

import numpy as np
import pytest

from gnnagg import synth


def _arr(x, dt=np.int32):
    return np.asarray(x, dt)


def test_oracle_schedules_match_golden(orc, golden):
    checked = 0
    for case in golden["cases"]:
        ptr, idx, val = _arr(case["ptr"]), _arr(case["idx"]), _arr(case["val"], np.float32)
        for s in case["schedules"]:
            if s["kind"] == 1:
                p, i, t = orc.neighbor_grouping(ptr, idx, s["neighbor_num"])
                v = None
            else:
                p, i, t, v = orc.locality(ptr, idx, s["par_num"], s["total_num_v"], val, s.get("neighbor_num", 0))
            assert p.tolist() == s["ptr"], (case["name"], s["kind"])
            assert i.tolist() == s["idx"]
            assert t.tolist() == s["target"]
            if v is not None:
                assert v.tolist() == s["val"]
            checked += 1
    assert checked > 100


def test_oracle_reorder_matches_golden(orc, golden):
    for case in golden["cases"]:
        r = case["reorder"]
        newptr, newidx = orc.reorder_csr(_arr(case["ptr"]), _arr(case["idx"]), _arr(r["rows"]), _arr(r["reverse_rows"]))
        assert newptr.tolist() == r["newptr"] and newidx.tolist() == r["newidx"]


def test_oracle_survey_vectors(orc):
    """the hand-checked vectors of SURVEY.md section 4"""
    ptr, idx = _arr([0, 3, 3, 8, 9]), _arr([1, 2, 3, 0, 1, 2, 3, 0, 2])
    val = np.arange(1, 10, dtype=np.float32)
    p, i, t = orc.neighbor_grouping(ptr, idx, 2)
    assert p.tolist() == [0, 2, 3, 5, 7, 8, 9] and t.tolist() == [0, 0, 2, 2, 2, 3] and i.tolist() == idx.tolist()
    p, i, t, v = orc.locality(ptr, idx, 2, 4, val, neighbor_num=2)
    assert p.tolist() == [0, 1, 3, 4, 6, 8, 9] and i.tolist() == [1, 0, 1, 0, 2, 3, 2, 3, 2]
    assert t.tolist() == [0, 2, 2, 0, 2, 3] and v.tolist() == [1, 4, 5, 8, 2, 3, 6, 7, 9]
    p, i, t, v = orc.locality(ptr, idx, 2, 4, val)
    assert p.tolist() == [0, 1, 4, 6, 8, 9] and t.tolist() == [0, 2, 0, 2, 3]
    p, i, t, _ = orc.locality(ptr, idx, 3, 4)
    assert p.tolist() == [0, 2, 3, 4, 6, 8, 9] and i.tolist() == [0, 0, 1, 1, 2, 3, 2, 3, 2] and t.tolist() == [2, 0, 2, 0, 2, 3]
    newptr, newidx = orc.reorder_csr(ptr, idx, _arr([2, 0, 3, 1]), _arr([1, 3, 0, 2]))
    assert newptr.tolist() == [0, 5, 8, 9, 9] and newidx.tolist() == [1, 3, 0, 2, 1, 3, 0, 2, 0]


def test_oracle_loader_matches_golden(orc, golden, tmp_path):
    g = golden["load_graph"]
    d = str(tmp_path) + "/"
    (tmp_path / "tiny.config").write_text(g["config"])
    (tmp_path / "tiny.graph").write_text(" ".join(map(str, g["graph_ptr"])) + "\n" + " ".join(map(str, g["graph_idx"])) + "\n")
    (tmp_path / "tiny.reorder_t").write_text(g["reorder_text"])
    ptr, idx, rows, rev = orc.load_graph(d, "tiny", d + "tiny.reorder_t")
    assert ptr.tolist() == g["reordered_ptr"] and idx.tolist() == g["reordered_idx"]
    assert rows.tolist() == g["rows"] and rev.tolist() == g["reverse_rows"]
    assert np.fromfile(d + "tiny.graph.ptrdump", np.int32).tolist() == g["ptrdump"]
    assert np.fromfile(d + "tiny.graph.edgedump", np.int32).tolist() == g["edgedump"]
    ptr, idx, rows, _ = orc.load_graph(d, "tiny")  # served from the caches
    assert ptr.tolist() == g["cached_ptr"] and idx.tolist() == g["cached_idx"] and rows is None


def test_oracle_float_semantics_tiny(orc):
    """hand-computed answers for the float restatements"""
    ptr, idx = _arr([0, 2, 2, 3]), _arr([1, 2, 0])
    val = _arr([0.5, 2.0, 3.0], np.float32)
    X = _arr([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], np.float32)
    Y = orc.spmm_f32(ptr, idx, val, X)
    assert Y.tolist() == [[20.5, 23.0, 25.5, 28.0], [0, 0, 0, 0], [3, 6, 9, 12]]
    Y64, S = orc.spmm_f64(ptr, idx, val, X)
    assert np.array_equal(Y, Y64) and np.array_equal(S, np.abs(Y64))
    att = _arr([[0.1, 0.2], [0.3, -0.4], [-2.0, 0.5]], np.float32)
    w01 = np.exp(max(0.1 - 0.4, 0.2 * (0.1 - 0.4)))
    w02 = np.exp(0.1 + 0.5)
    G, den, _ = orc.gat_f64(ptr, idx, att, X)
    np.testing.assert_allclose(den[0], w01 + w02, rtol=1e-6)
    np.testing.assert_allclose(G[0], (w01 * X[1] + w02 * X[2]) / (w01 + w02), rtol=1e-6)
    assert np.all(G[1] == 0) and np.allclose(G[2], X[0])
    Gn, _, _ = orc.gat_f64(ptr, idx, att, X, empty_value=float("nan"))
    assert np.all(np.isnan(Gn[1]))  # the reference's 0/0 (aggr_gat.h:163)
    sm = orc.edge_softmax_f64(ptr, idx, att)
    np.testing.assert_allclose(sm, [w01 / (w01 + w02), w02 / (w01 + w02), 1.0], rtol=1e-6)
    uv = orc.u_add_v(ptr, idx, att)
    np.testing.assert_allclose(uv, [0.1 - 0.4, 0.1 + 0.5, -2.0 + 0.2], rtol=1e-6)
    center = orc.add_to_center_f64(ptr, uv)
    np.testing.assert_allclose(center, [uv[0] + uv[1], 0.0, uv[2]], rtol=1e-6)
    sd, _ = orc.sddmm_f64(ptr, idx, X, X)
    assert sd.tolist() == [float(X[1] @ X[0]), float(X[2] @ X[0]), float(X[0] @ X[2])]
    W = _arr([[1, 0], [0, 1], [1, 1], [2, -1]], np.float32)
    AX, H, _ = orc.gcn_layer_f64(ptr, idx, val, X, W)
    assert np.array_equal(AX, Y) and np.array_equal(H, Y @ W)
    el = orc.csr2edgelist(ptr, idx)
    assert el.tolist() == [1, 0, 2, 0, 0, 2]
    gp, gi, gt = orc.neighbor_grouping(ptr, idx, 1)
    Yg, _ = orc.spmm_grouped_f64(3, gp, gi, val, gt, X)
    assert np.array_equal(Yg, Y)


@pytest.mark.parametrize("seed", range(6))
def test_oracle_equals_compiled_reference_random(orc, seed):
    """oracle vs the reference's host code compiled from /root/reference (oracle/_ref); runs wherever
    libref.so exists (it travels to the GPU box), skipped otherwise"""
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    ptr, idx = synth.small_random_csr(n, float(rng.uniform(0.5, 20)), seed, empty_frac=float(rng.uniform(0, 0.6)),
                                      hub=int(rng.integers(0, 2) * rng.integers(50, 600)))
    val = rng.standard_normal(len(idx)).astype(np.float32)
    for ng in (1, 7, 32):
        a = orc.neighbor_grouping(ptr, idx, ng)
        b = orc.ref_schedule(1, ptr, idx, neighbor_num=ng)
        assert all(np.array_equal(x, y) for x, y in zip(a, b[:3]))
    for par in (1, 2, 5):
        for total in (n, max(1, n // 2)):
            a = orc.locality(ptr, idx, par, total, val)
            b = orc.ref_schedule(0, ptr, idx, val, par_num=par, total_num_v=total)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            a = orc.locality(ptr, idx, par, total, val, neighbor_num=5)
            b = orc.ref_schedule(2, ptr, idx, val, par_num=par, neighbor_num=5, total_num_v=total)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    rows = rng.permutation(n).astype(np.int32)
    rev = np.empty(n, np.int32)
    rev[rows] = np.arange(n, dtype=np.int32)
    a, b = orc.reorder_csr(ptr, idx, rows, rev), orc.ref_reorder_csr(ptr, idx, rows, rev)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
