"""helpers shared by the GPU parity tests"""
import numpy as np
import torch

from gnnagg import synth


def graph(n, avg, seed, empty_frac=0.2, hub=0):
    ptr, idx = synth.small_random_csr(n, avg, seed, empty_frac=empty_frac, hub=hub)
    return ptr, idx


def dev(a, device="cuda:0"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def rand_inputs(n, m, F, seed, positive=False):
    rng = np.random.default_rng(seed)
    if positive:  # the all-positive parity set of SURVEY 8(d): no cancellation, plain relative error is meaningful
        X = rng.random((n, F), dtype=np.float32)
        val = (rng.random(m, dtype=np.float32) + 0.1).astype(np.float32)
    else:         # what the reference drivers feed: N(0,1) features and edge values (Figure9/main.cu:44-50)
        X = rng.standard_normal((n, F)).astype(np.float32)
        val = rng.standard_normal(m).astype(np.float32)
    return X, val


# (n, avg_deg, empty_frac, hub) -- edge cases the walk must survive
GRAPHS = {
    "tiny": (5, 2.0, 0.3, 0),
    "one_row": (1, 40.0, 0.0, 0),
    "short_rows": (3000, 3.0, 0.3, 0),
    "medium": (700, 40.0, 0.1, 0),
    "hub": (300, 8.0, 0.2, 9000),          # one row far longer than an item
    "bighub": (60, 4.0, 0.2, 40000),       # spans > 32 items of 512 edges: two-pass fix-up
    "leading_trailing_empty": (-1, 0, 0, 0),  # built by hand below
    "exact_items": (-2, 0, 0, 0),
}


def make_graph(name, seed=0):
    if name == "leading_trailing_empty":
        deg = np.array([0, 0, 0, 5, 0, 513, 0, 0, 1200, 0, 0, 0], np.int64)
    elif name == "exact_items":
        deg = np.array([512, 512, 0, 1024, 128, 128, 256], np.int64)  # rows end exactly on item boundaries
    else:
        n, avg, ef, hub = GRAPHS[name]
        return graph(n, avg, seed, ef, hub)
    ptr = np.zeros(len(deg) + 1, np.int32)
    ptr[1:] = np.cumsum(deg)
    idx = np.random.default_rng(seed).integers(0, len(deg), int(ptr[-1])).astype(np.int32)
    return ptr, idx
