"""GPU tier: gnnagg.plugin -- the reference's PyTorch extension surface (Figure7/kernel.cpp:37-179) -- used the
way Figure7/our.py uses `gnncompile`: new_load -> gcn_init/gat_init -> *_schedule(32) -> 3 layers of
torch.mm + aggregation (512 -> 128 -> 64 -> 32, our.py:84-87,171-188)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_gate
from gnnagg import synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gnc(gn, cuda, tmp_path, monkeypatch):
    import gnnagg.plugin as plugin

    (tmp_path / "data").mkdir()
    (tmp_path / "run").mkdir()
    ptr, idx = synth.rmat_csr(3000, 40000, seed=9)
    gn.write_graph("syn", ptr.numpy(), idx.numpy(), str(tmp_path / "data") + "/")
    rows = np.random.default_rng(1).permutation(3000).astype(np.int32)
    gn.write_reorder(str(tmp_path / "data" / "syn.reorder_thres_0.2"), rows)
    monkeypatch.chdir(tmp_path / "run")  # the reference reads ../data relative to the CWD (src/data.cu:34)
    return plugin


def test_three_layer_gcn_and_gat_like_our_py(gnc, orc, cuda):
    ptrs, idxs = gnc.new_load("syn", "_thres_0.2", 0)
    assert ptrs.dtype == torch.int32 and ptrs.is_cuda and gnc.n == 3000 and gnc.m == 40000
    num_v, num_e = gnc.n, gnc.m
    vals = torch.ones(num_e, device=cuda)                     # our.py:78
    at = gnc.gcn_init(ptrs, idxs, vals)
    gnc.gcn_schedule(at, 32)
    at_gat = gnc.gat_init(ptrs, idxs)
    gnc.gat_schedule(at_gat, 32)
    torch.manual_seed(123)
    dims = [512, 128, 64, 32]
    weights = [torch.randn(dims[i], dims[i + 1], device=cuda) / dims[i] ** 0.5 for i in range(3)]
    h0 = torch.randn(num_v, 512, device=cuda)
    hp, hi = ptrs.cpu().numpy(), idxs.cpu().numpy()

    # GCN: feat2 = mm(feat, W); gcn_run(at, feat2, out, 128, 1); relu          (our.py:171-176)
    h, h_ref = h0, h0.cpu().numpy()
    for w in weights:
        feat2 = torch.mm(h, w)
        out = torch.empty_like(feat2)
        gnc.gcn_run(at, feat2, out, 128, 1)
        y64, scale = orc.spmm_f64(hp, hi, vals.cpu().numpy(), feat2.cpu().numpy())
        assert rel_gate(out.cpu().numpy(), y64, scale, 1e-5)[0] == 0
        h = torch.relu(out)

    # GAT: att_lr = mm(feat2, W_lr); gat_run(at_gat, feat2, att_lr, out, 128, 1)   (our.py:178-188)
    h = h0
    for w in weights:
        feat2 = torch.mm(h, w)
        att_lr = torch.mm(feat2, torch.randn(feat2.shape[1], 2, device=cuda) / feat2.shape[1] ** 0.5)
        out = torch.empty_like(feat2)
        gnc.gat_run(at_gat, feat2, att_lr, out, 128, 1)
        y64, _, scale = orc.gat_f64(hp, hi, att_lr.cpu().numpy(), feat2.cpu().numpy())
        assert rel_gate(out.cpu().numpy(), y64, scale, 1.2e-5)[0] == 0
        h = out

    # the un-fused pieces the script keeps commented (our.py:160-168) and gcn_update_val
    att = torch.randn(num_v, 2, device=cuda)
    val_mid = torch.empty(num_e, device=cuda)
    gnc.gat_run_u_add_v(at_gat, att, val_mid, 128)
    val_mid = torch.exp(torch.nn.functional.leaky_relu(val_mid, 0.2))
    att_mid = torch.empty(num_v, device=cuda)
    gnc.gat_run_add_to_center(at_gat, val_mid, att_mid, 128)
    gnc.gat_run_div_each(at_gat, att_mid, val_mid, 128)
    np.testing.assert_allclose(val_mid.cpu().numpy(), orc.edge_softmax_f64(hp, hi, att.cpu().numpy()), rtol=2e-5, atol=1e-12)
    gnc.gcn_update_val(at, val_mid)
    feat = torch.randn(num_v, 32, device=cuda)
    out = torch.empty_like(feat)
    gnc.gcn_run(at, feat, out, 128, 1)
    y64, _, scale = orc.gat_f64(hp, hi, att.cpu().numpy(), feat.cpu().numpy())
    assert rel_gate(out.cpu().numpy(), y64, scale, 3e-5)[0] == 0
    gnc.destroy(at)
    gnc.destroy(at_gat)
    with pytest.raises(Exception):
        gnc.gcn_run(at, feat, out, 128, 0)


def test_plugin_input_checks(gnc, cuda):
    ptrs, idxs = gnc.new_load("syn", "", 0)
    with pytest.raises(Exception):
        gnc.gcn_init(ptrs.cpu(), idxs, torch.ones(gnc.m, device=cuda))     # CHECK_CUDA
    at = gnc.gcn_init(ptrs, idxs, torch.ones(gnc.m, device=cuda))
    x = torch.randn(gnc.n, 64, device=cuda)
    with pytest.raises(Exception):
        gnc.gcn_run(at, x.t()[:32].t(), torch.empty(gnc.n, 32, device=cuda), 128, 0)  # CHECK_CONTIGUOUS


def test_autograd_functions_match_the_oracle_gradients(gnc, orc, cuda):
    """gnc.gcn_aggregate / gnc.gat_aggregate under torch.autograd: forward against the fp64 oracle, backward against
    the oracle's A^T dY and full GAT derivative (oracle pinned by finite differences in tests/test_backward.py)"""
    from gnnagg import synth

    ptr, idx = synth.small_random_csr(900, 12.0, 31, hub=3000)
    rng = np.random.default_rng(4)
    n, F = len(ptr) - 1, 32
    val = (rng.random(len(idx)).astype(np.float32) + 0.1)
    X = rng.standard_normal((n, F)).astype(np.float32)
    att = rng.standard_normal((n, 2)).astype(np.float32)
    dY = rng.standard_normal((n, F)).astype(np.float32)
    dptr, didx, dval = (torch.from_numpy(a).to(cuda) for a in (ptr, idx, val))
    at = gnc.gcn_init(dptr, didx, dval)
    Xt = torch.from_numpy(X).to(cuda).requires_grad_(True)
    Y = gnc.gcn_aggregate(at, Xt)
    Y.backward(torch.from_numpy(dY).to(cuda))
    y64, scale = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(Y.detach().cpu().numpy(), y64, scale, 1e-5)[0] == 0
    dx64, dscale = orc.spmm_t_f64(ptr, idx, val, dY, n)
    assert rel_gate(Xt.grad.cpu().numpy(), dx64, dscale, 1e-5)[0] == 0

    at_gat = gnc.gat_init(dptr, didx)
    Xg = torch.from_numpy(X).to(cuda).requires_grad_(True)
    ag = torch.from_numpy(att).to(cuda).requires_grad_(True)
    Yg = gnc.gat_aggregate(at_gat, Xg, ag)
    Yg.backward(torch.from_numpy(dY).to(cuda))
    g64, _, gs = orc.gat_f64(ptr, idx, att, X)
    assert rel_gate(Yg.detach().cpu().numpy(), g64, gs, 1.2e-5)[0] == 0
    ref = orc.gat_backward_f64(ptr, idx, att, X, dY)
    dX64, datt64 = ref[0], ref[1]
    assert np.abs(Xg.grad.cpu().numpy() - dX64).max() <= 5e-5 * max(1.0, np.abs(dX64).max())
    assert np.abs(ag.grad.cpu().numpy() - datt64).max() <= 1e-4 * max(1.0, np.abs(datt64).max())
    gnc.destroy(at)
    gnc.destroy(at_gat)
