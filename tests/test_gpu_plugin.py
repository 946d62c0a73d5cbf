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
