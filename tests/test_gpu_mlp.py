"""GPU tier: the per-edge MLP aggregator (include/aggr_nn.h of the reference).
Y[v] = sum_u ReLU((X[v] + X[u]) W).  Ours hoists the projection (P = X W on tcgen05, then gather + ReLU + sum);
checked against the fp64 restatement of the reference's per-edge mat-vec and against the reference's own
aggr_mlp / aggr_mlp_target kernels (F = 32) recompiled for sm_100."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_gate
from gpu_util import dev, make_graph

pytestmark = pytest.mark.gpu


def _inputs(n, F, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, F)).astype(np.float32), (rng.standard_normal((F, F)) / np.sqrt(F)).astype(np.float32)


@pytest.mark.parametrize("gname", ["tiny", "short_rows", "medium", "hub", "leading_trailing_empty", "exact_items"])
@pytest.mark.parametrize("F", [32, 64, 128])
def test_mlp_parity(gn, orc, cuda, gname, F):
    ptr, idx = make_graph(gname, seed=F + 3)
    n = len(ptr) - 1
    X, W = _inputs(n, F, 71)
    agg = gn.Aggregator(dev(ptr), dev(idx))
    Y = agg.mlp_run(dev(X), dev(W), torch.full((n, F), float("nan"), device=cuda))
    y64, scale = orc.mlp_f64(ptr, idx, X, W)
    bad, worst = rel_gate(Y.cpu().numpy(), y64, scale, 1e-5)
    assert bad == 0, (gname, F, worst)
    assert torch.equal(Y, agg.mlp_run(dev(X), dev(W), torch.empty((n, F), device=cuda)))  # deterministic
    agg.schedule(1, [16])
    Ys = agg.mlp_run(dev(X), dev(W), torch.full((n, F), float("nan"), device=cuda), scheduled=True)
    assert rel_gate(Ys.cpu().numpy(), y64, scale, 1e-5)[0] == 0


def test_mlp_vs_reference_kernels(gn, orc, cuda):
    if not orc.ref_available():
        pytest.skip("oracle/_ref/libref.so not present")
    ref = orc.ref()
    ptr, idx = make_graph("medium", seed=12)
    n, m = len(ptr) - 1, len(idx)
    X, W = _inputs(n, 32, 72)
    dptr, didx, dX, dW = dev(ptr), dev(idx), dev(X), dev(W)
    P = lambda t: C.c_void_p(t.data_ptr())
    ref.ref_set_globals(n, m)
    h = C.c_void_p(ref.ref_mlp_create(P(dptr), P(didx), n, m, P(dW)))
    Yr = torch.zeros((n, 32), device=cuda)
    ref.ref_mlp_run(h, P(dX), P(Yr), 128, 0)              # aggr_mlp
    ref.ref_mlp_schedule(h, 1, 16, 0)
    Yr2 = torch.zeros((n, 32), device=cuda)
    ref.ref_mlp_run(h, P(dX), P(Yr2), 128, 1)             # aggr_mlp_target
    agg = gn.Aggregator(dptr, didx)
    Y = agg.mlp_run(dX, dW, torch.empty((n, 32), device=cuda))
    y64, scale = orc.mlp_f64(ptr, idx, X, W)
    assert rel_gate(Y.cpu().numpy(), y64, scale, 1e-5)[0] == 0
    for theirs in (Yr, Yr2):
        _, worst_r = rel_gate(theirs.cpu().numpy(), y64, scale, 1e-5)
        print("reference aggr_mlp worst err/bound %.3f" % worst_r)
        assert rel_gate(Y.cpu().numpy(), theirs.cpu().numpy(), scale, 1e-5 * (1 + max(worst_r, 1.0)))[0] == 0
