#!/usr/bin/env python
"""A few fused-GAT aggregations on the proteins-shaped graph (C3, F=64): the command the GAT
ncu capture of tools/profile_round.sh profiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))
import torch

import gnnagg
from gnnagg import synth

dev = torch.device("cuda:0")
n, m, F = synth.SHAPES["proteins"][:2] + (64,)
ptr, idx = synth.rmat_csr(n, m, seed=123, device=dev)
g = torch.Generator(device=dev).manual_seed(123)
X = torch.randn((n, F), device=dev, generator=g)
att = torch.randn((n, 2), device=dev, generator=g)
agg = gnnagg.Aggregator(ptr, idx, None)
Y = torch.empty((n, F), device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    agg.gat_run(X, att, Y)
torch.cuda.synchronize()
print("gat_loop ok", float(Y.abs().mean()))
