"""GPU tier: the dense combination on tcgen05 (3xTF32) and the fused GCN layer H = (A X) W."""
import numpy as np
import pytest
import torch

from conftest import rel_gate
from gpu_util import dev, make_graph, rand_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (1, 128, 128), (300, 128, 128), (1000, 64, 256), (777, 256, 64),
                                   (513, 256, 256), (4096, 128, 32), (129, 96, 160), (4097, 256, 128), (300, 128, 256),
                                   (70000, 256, 256), (1, 192, 224), (100000, 128, 128), (60000, 32, 32), (50000, 64, 224)])
def test_dense_nn_parity(gn, orc, cuda, M, K, N):
    rng = np.random.default_rng(M + K + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    C = gn.dense_nn(dev(A), dev(W), torch.full((M, N), float("nan"), device=cuda))
    torch.cuda.synchronize()
    h64, scale = orc.dense_f64(A, W)
    bad, worst = rel_gate(C.cpu().numpy(), h64, scale, TOL)
    assert bad == 0, worst


@pytest.mark.parametrize("M,K,N", [(8, 48, 32), (1000, 100, 7), (333, 16, 16), (257, 320, 64), (1, 1, 1), (5000, 40, 300)])
def test_dense_nn_generic_shapes(gn, orc, cuda, M, K, N):
    """shapes the tensor-core kernel does not take (K or N not a multiple of 32, or above 256) go through the plain fp32
    kernel, as any size goes through cuBLAS in the reference's matmul_NN (dense.h:4-23)"""
    rng = np.random.default_rng(M + K + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    C = gn.dense_nn(dev(A), dev(W), torch.full((M, N), float("nan"), device=cuda))
    h64, scale = orc.dense_f64(A, W)
    bad, worst = rel_gate(C.cpu().numpy(), h64, scale, TOL)
    assert bad == 0, worst


@pytest.mark.parametrize("fin,fout", [(128, 128), (32, 32), (64, 32), (256, 128)])
@pytest.mark.parametrize("scheduled", [False, True])
def test_gcn_layer_parity(gn, orc, cuda, fin, fout, scheduled):
    """Figure10/main_b.cu: aggregation + combination (run_with_nn vs run + matmul_NN)"""
    ptr, idx = make_graph("hub", seed=fin + fout)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, fin, seed=11)
    W = (np.random.default_rng(3).standard_normal((fin, fout)) / np.sqrt(fin)).astype(np.float32)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    if scheduled:
        agg.schedule(1, [64])  # Figure10/run.sh:14 --nei 64
    AX = torch.full((n, fin), float("nan"), device=cuda)
    H = agg.gcn_layer(dev(X), dev(W), torch.full((n, fout), float("nan"), device=cuda), AX, scheduled=scheduled)
    ax64, h64, scale = orc.gcn_layer_f64(ptr, idx, val, X, W)
    _, ax_scale = orc.spmm_f64(ptr, idx, val, X)
    assert rel_gate(AX.cpu().numpy(), ax64, ax_scale, TOL)[0] == 0
    bad, worst = rel_gate(H.cpu().numpy(), h64, scale, TOL)
    assert bad == 0, worst
    # AX may be omitted
    H2 = agg.gcn_layer(dev(X), dev(W), torch.empty((n, fout), device=cuda), None, scheduled=scheduled)
    assert rel_gate(H2.cpu().numpy(), h64, scale, TOL)[0] == 0


def test_gcn_layer_host_entry_point(gn, orc, cuda):
    ptr, idx = make_graph("medium", seed=21)
    n, m = len(ptr) - 1, len(idx)
    X, val = rand_inputs(n, m, 128, seed=12)
    W = (np.random.default_rng(5).standard_normal((128, 128)) / np.sqrt(128)).astype(np.float32)
    agg = gn.Aggregator(dev(ptr), dev(idx), dev(val))
    hX, hW = torch.from_numpy(X).pin_memory(), torch.from_numpy(W).pin_memory()
    hH = torch.empty((n, 128)).pin_memory()
    agg.gcn_layer_host(hX, hW, hH)
    _, h64, scale = orc.gcn_layer_f64(ptr, idx, val, X, W)
    assert rel_gate(hH.numpy(), h64, scale, TOL)[0] == 0
