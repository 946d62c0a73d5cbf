"""Synthetic power-law graphs of the reference's dataset shapes.

The dataset download of the reference (README.md:45-54) is not available offline, so inputs are
R-MAT graphs (a=0.57 b=0.19 c=0.19 d=0.05) with exactly the vertex/edge counts of the named
datasets (util.py:24-142 of the reference).  Every edge is a pure function of (seed, edge id), built
from splitmix64 with integer tensor ops only, so the same graph comes out of torch on a CPU or on
any GPU.  Vertex ids are scrambled by an odd-multiplier bijection so id order carries no locality;
endpoints >= n are rejected and redrawn; duplicates and self loops are kept; edges are sorted by
(dst, src) into CSR with row = destination.
"""
import numpy as np
import torch

SHAPES = {
    # name: (num_v, num_e) -- BASELINE.json configs / reference util.py
    "arxiv": (169343, 1166243),
    "reddit": (232965, 114615891),
    "proteins": (132534, 39561252),
    "products": (2449029, 61859140),
    "ddi": (4267, 2135822),
    "collab": (235868, 2358104),
}

_A, _B, _C = 0.57, 0.19, 0.19
_MASK63 = (1 << 63) - 1


def _i64(v):
    """python int -> wrapped int64 constant"""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, k):
    """logical shift right of an int64 tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x):
    x = x + _i64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _rmat_pairs(eid, attempt, seed, scale):
    """(row bits, col bits) of `scale` levels for edge ids `eid` (int64 tensor)"""
    ta, tb, tc = int(_A * (1 << 24)), int((_A + _B) * (1 << 24)), int((_A + _B + _C) * (1 << 24))
    base = splitmix64(eid * _i64(0xD1342543DE82EF95) + _i64(seed * 0x2545F4914F6CDD1D + attempt * 0x632BE59BD9B4E019))
    r = torch.zeros_like(eid)
    c = torch.zeros_like(eid)
    for level in range(scale):
        h = splitmix64(base + _i64(level * 0x9E3779B97F4A7C15))
        u = _lsr(h, 40)  # 24 uniform bits
        # quadrant: a -> (0,0), b -> (0,1), c -> (1,0), d -> (1,1)
        rb = (u >= tb).to(torch.int64)
        cb = (((u >= ta) & (u < tb)) | (u >= tc)).to(torch.int64)
        r = (r << 1) | rb
        c = (c << 1) | cb
    return r, c


def _scramble(v, scale, mul):
    return (v * mul + (mul >> 3)) & ((1 << scale) - 1)


def rmat_edges(num_v, num_e, seed=123, device="cpu", chunk=1 << 24, src_num_v=None, dst_prefix=None):
    """dst, src int64 tensors of `num_e` edges with dst in [0,num_v), src in [0,src_num_v or num_v)"""
    src_n = num_v if src_num_v is None else src_num_v
    sd = max(1, int(np.ceil(np.log2(max(num_v, 2)))))
    ss = max(1, int(np.ceil(np.log2(max(src_n, 2)))))
    scale = max(sd, ss)
    dst_all = torch.empty(num_e, dtype=torch.int64, device=device)
    src_all = torch.empty(num_e, dtype=torch.int64, device=device)
    off = 0 if dst_prefix is None else int(dst_prefix) * (1 << 40)
    for lo in range(0, num_e, chunk):
        hi = min(num_e, lo + chunk)
        eid = torch.arange(lo, hi, dtype=torch.int64, device=device) + off
        dst = torch.empty(hi - lo, dtype=torch.int64, device=device)
        src = torch.empty(hi - lo, dtype=torch.int64, device=device)
        todo = torch.arange(hi - lo, device=device)
        attempt = 0
        while todo.numel() > 0:
            r, c = _rmat_pairs(eid[todo], attempt, seed, scale)
            d = _scramble(r >> (scale - sd), sd, 0x9E3779B1)
            s = _scramble(c >> (scale - ss), ss, 0x85EBCA77)
            ok = (d < num_v) & (s < src_n)
            dst[todo[ok]] = d[ok]
            src[todo[ok]] = s[ok]
            todo = todo[~ok]
            attempt += 1
        dst_all[lo:hi] = dst
        src_all[lo:hi] = src
    return dst_all, src_all


def to_csr(dst, src, num_v, src_num_v=None):
    """sort by (dst, src) -> int32 ptr[num_v+1], idx[num_e]"""
    src_n = num_v if src_num_v is None else src_num_v
    bits = max(1, int(np.ceil(np.log2(max(src_n, 2)))))
    key = (dst << bits) | src
    key, _ = torch.sort(key)
    idx = (key & ((1 << bits) - 1)).to(torch.int32)
    rows = key >> bits
    counts = torch.bincount(rows, minlength=num_v)
    ptr = torch.zeros(num_v + 1, dtype=torch.int64, device=dst.device)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr.to(torch.int32), idx


def rmat_csr(num_v, num_e, seed=123, device="cpu", src_num_v=None, dst_prefix=None):
    dst, src = rmat_edges(num_v, num_e, seed, device, src_num_v=src_num_v, dst_prefix=dst_prefix)
    return to_csr(dst, src, num_v, src_num_v)


def gcn_norm_val(ptr, idx, src_deg=None):
    """val = 1/sqrt((deg_dst+1)(deg_src+1)) with deg = CSR row length (GCN normalisation, positive)"""
    deg = (ptr[1:] - ptr[:-1]).to(torch.float32)
    sdeg = deg if src_deg is None else src_deg
    row = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=ptr.device), (ptr[1:] - ptr[:-1]).long())
    return torch.rsqrt((deg[row] + 1.0) * (sdeg[idx.long()] + 1.0))


def shape_of(name):
    return SHAPES[name]


def small_random_csr(num_v, avg_deg, seed, empty_frac=0.2, hub=0, num_src=None):
    """numpy CSR for tests: random rows, a fraction of empty rows, optional hub row, duplicates kept"""
    rng = np.random.default_rng(seed)
    num_src = num_v if num_src is None else num_src
    deg = rng.poisson(avg_deg, num_v).astype(np.int64)
    deg[rng.random(num_v) < empty_frac] = 0
    if hub and num_v > 0:
        deg[rng.integers(num_v)] = hub
    ptr = np.zeros(num_v + 1, np.int32)
    ptr[1:] = np.cumsum(deg)
    idx = rng.integers(0, max(num_src, 1), int(ptr[-1])).astype(np.int32)
    return ptr, idx


def planted_community_csr(num_v, comm=48, pool=48, picks=24, noise=4, seed=123, device="cpu"):
    """graph WITH community structure and scrambled ids: the `comm` members of a community draw `picks`
    neighbours from a community-specific pool of `pool` random vertices plus `noise` uniformly random
    ones.  Members of one community have Jaccard similarity ~0.25 (above the 0.2 LSH threshold of
    script/cluster2.py) but are scattered over the id range, which is the situation the reference's
    locality reorder is built for (R-MAT graphs have no such structure to recover)."""
    g = torch.Generator(device=device).manual_seed(seed)
    ncomm = (num_v + comm - 1) // comm
    perm = torch.randperm(num_v, generator=g, device=device)               # vertex -> scrambled id
    pools = torch.randint(0, num_v, (ncomm, pool), generator=g, device=device)
    member_comm = torch.arange(num_v, device=device) // comm
    sel = torch.rand((num_v, pool), generator=g, device=device).argsort(1)[:, :picks]   # picks distinct pool slots
    nbr = torch.gather(pools[member_comm], 1, sel)
    rnd = torch.randint(0, num_v, (num_v, noise), generator=g, device=device)
    nbr = torch.cat([nbr, rnd], 1)                                          # [num_v, picks+noise]
    dst = perm.repeat_interleave(picks + noise)
    return to_csr(dst, nbr.reshape(-1), num_v)
