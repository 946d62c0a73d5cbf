#!/usr/bin/env python
"""A few GAT backward passes on a BASELINE.json shape: the command behind the backward launch lists in profiles/.
usage: bwd_loop.py [shape] [F] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))
import torch

import gnnagg
from gnnagg import synth

shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
n, m = synth.shape_of(shape)
ptr, idx = synth.rmat_csr(n, m, seed=123, device=dev)
g = torch.Generator(device=dev).manual_seed(123)
X = torch.randn((n, F), device=dev, generator=g)
dY = torch.randn((n, F), device=dev, generator=g)
att = torch.randn((n, 2), device=dev, generator=g)
agg = gnnagg.Aggregator(ptr, idx, None)
agg.transpose_build()
Y = torch.empty((n, F), device=dev)
dX = torch.empty((n, F), device=dev)
dA = torch.empty((n, 2), device=dev)
agg.gat_run(X, att, Y)
for _ in range(reps):
    agg.gat_backward(X, att, Y, dY, dX, dA)
torch.cuda.synchronize()
print("bwd_loop ok", float(dX.abs().mean()), float(dA.abs().mean()))
