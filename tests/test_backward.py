"""CPU tier: the backward oracle (oracle.c: orc_transpose_csr, orc_spmm_t_f64, orc_gat_backward_f64) is
pinned against finite differences of the forward oracle -- the reference has no correct backward to
take golden vectors from (aggr_gat.h:222-294 is experimental: F = 32 only, LeakyReLU derivative keyed
on `newval < 0`, source half of the attention gradient only); the GPU tier compares with that kernel
where it is right (tests/test_gpu_backward.py)."""
import numpy as np
import pytest

from gnnagg import synth


def graph(n, deg, seed, num_src=None, hub=0):
    ptr, idx = synth.small_random_csr(n, deg, seed, empty_frac=0.2, hub=hub, num_src=num_src)
    return ptr.astype(np.int32), idx.astype(np.int32)


@pytest.mark.parametrize("n,num_src,deg", [(1, 1, 3), (50, 50, 4), (64, 200, 7), (200, 40, 5), (30, 30, 0)])
def test_transpose_is_stable_sort_by_source(orc, n, num_src, deg):
    ptr, idx = graph(n, deg, 7, num_src)
    t_ptr, t_idx, t_perm = orc.transpose_csr(ptr, idx, num_src)
    m = len(idx)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(ptr))
    order = np.argsort(idx, kind="stable").astype(np.int32)
    assert np.array_equal(t_perm, order)
    assert np.array_equal(t_idx, rows[order])
    assert np.array_equal(t_ptr, np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=num_src))]).astype(np.int32))
    assert t_ptr[-1] == m


def test_spmm_t_is_the_adjoint(orc):
    """<A X, dY> == <X, A^T dY> for every X, dY: fp64 accumulation on both sides"""
    n, num_src, F = 120, 90, 8
    ptr, idx = graph(n, 6, 3, num_src, hub=40)
    rng = np.random.default_rng(0)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((num_src, F)).astype(np.float32)
    dY = rng.standard_normal((n, F)).astype(np.float32)
    Y, _ = orc.spmm_f64(ptr, idx, val, X)
    dX, S = orc.spmm_t_f64(ptr, idx, val, dY, num_src)
    lhs = float((Y.astype(np.float64) * dY).sum())
    rhs = float((X.astype(np.float64) * dX).sum())
    assert abs(lhs - rhs) <= 1e-5 * float((np.abs(X) * S).sum())
    # and element-wise against a dense product
    A = np.zeros((n, num_src))
    np.add.at(A, (np.repeat(np.arange(n), np.diff(ptr)), idx), val.astype(np.float64))
    assert np.allclose(dX, A.T @ dY.astype(np.float64), rtol=0, atol=1e-5 * S.max())


@pytest.mark.parametrize("slope", [0.2, 1.0, 0.0])
@pytest.mark.parametrize("n,num_src,F", [(40, 40, 4), (25, 60, 8)])
def test_gat_backward_matches_finite_differences(orc, n, num_src, F, slope):
    ptr, idx = graph(n, 5, 11, num_src, hub=min(n, num_src) // 2)
    rows = max(n, num_src)
    rng = np.random.default_rng(1)
    att = rng.standard_normal((rows, 2)).astype(np.float32)
    # keep every pre-activation away from the LeakyReLU kink so central differences are valid
    dst = np.repeat(np.arange(n), np.diff(ptr))
    for _ in range(50):
        s = att[dst, 0] + att[idx, 1]
        bad = np.abs(s) < 0.05
        if not bad.any():
            break
        att[idx[bad], 1] += 0.11
    assert not (np.abs(att[dst, 0] + att[idx, 1]) < 0.02).any()
    X = rng.standard_normal((num_src, F)).astype(np.float32)
    dY = rng.standard_normal((n, F)).astype(np.float32)
    dX, dA, SX, SA = orc.gat_backward_f64(ptr, idx, att, X, dY, slope)

    a64, x64, d64 = att.astype(np.float64), X.astype(np.float64), dY.astype(np.float64)
    h = 1e-5

    def fd(arr, pos):
        keep = arr[pos]
        arr[pos] = keep + h
        up = orc.gat_loss_f64(ptr, idx, a64, x64, d64, slope)
        arr[pos] = keep - h
        dn = orc.gat_loss_f64(ptr, idx, a64, x64, d64, slope)
        arr[pos] = keep
        return (up - dn) / (2 * h)

    for pos in [tuple(p) for p in np.stack([rng.integers(0, num_src, 40), rng.integers(0, F, 40)], 1)]:
        assert abs(fd(x64, pos) - dX[pos]) <= 1e-5 * (SX[pos] + 1.0), pos
    for pos in [tuple(p) for p in np.stack([rng.integers(0, rows, 60), rng.integers(0, 2, 60)], 1)]:
        assert abs(fd(a64, pos) - dA[pos]) <= 1e-5 * (SA[pos] + 1.0), pos
    # rows that are neither a destination nor a source get exactly 0
    used_dst = np.zeros(rows, bool)
    used_dst[:n] = np.diff(ptr) > 0
    used_src = np.zeros(rows, bool)
    used_src[idx] = True
    assert (dA[~used_dst, 0] == 0).all() and (dA[~used_src, 1] == 0).all()
    assert (dX[~used_src[:num_src]] == 0).all()
