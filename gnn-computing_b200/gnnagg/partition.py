"""partition.py -- 1-D row partitioning by destination for multi-GPU aggregation.

The reference is single-GPU (every driver asserts GPUNUM == 1, Figure9/main.cu:19; the prepare*Multi
prototypes of include/data.h:48-58 have no definitions), so this is new functionality built around the
same aggregator: rank p owns a contiguous block of destination rows and the matching X / Y shard; its
local CSR keeps GLOBAL source ids and gathers from a replicated X_full that an all-gather of the X
shards fills (NCCL over NVLink on GPUs, gloo in the CPU tests).  No partial sums cross ranks.
"""
import numpy as np


def split_rows(ptr, parts, balance="edges"):
    """row boundaries [parts+1] of contiguous blocks; balance='edges' puts ~m/parts edges in each block
    (power-law graphs make equal-row blocks unbalanced), 'rows' gives equal row counts"""
    ptr = np.asarray(ptr)
    n = len(ptr) - 1
    if balance == "rows":
        return np.array([(n * p) // parts for p in range(parts + 1)], np.int64)
    m = int(ptr[-1])
    cuts = [0]
    for p in range(1, parts):
        target = (m * p) // parts
        r = int(np.searchsorted(ptr, target, side="left"))
        cuts.append(min(max(r, cuts[-1]), n))
    cuts.append(n)
    return np.array(cuts, np.int64)


def local_block(ptr, idx, val, bounds, rank):
    """(ptr_local, idx_local, val_local) of rank's row block; source ids stay global"""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    e0, e1 = int(ptr[r0]), int(ptr[r1])
    lp = (np.asarray(ptr[r0:r1 + 1]) - e0).astype(np.int32)
    return lp, np.ascontiguousarray(idx[e0:e1]), (None if val is None else np.ascontiguousarray(val[e0:e1]))


def gather_sizes(bounds):
    """rows owned by every rank (the all-gather of X is ragged when blocks are edge-balanced)"""
    return [int(bounds[p + 1] - bounds[p]) for p in range(len(bounds) - 1)]


def all_gather_rows(x_shard, bounds, group=None):
    """X_full [n, F] from per-rank shards [rows_p, F]; works for equal and ragged blocks and for any
    torch.distributed backend (nccl on GPUs, gloo on CPU)"""
    import torch
    import torch.distributed as dist

    sizes = gather_sizes(bounds)
    F = x_shard.shape[1]
    full = torch.empty((int(bounds[-1]), F), dtype=x_shard.dtype, device=x_shard.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(full, x_shard.contiguous(), group=group)
    else:
        # ragged blocks: one broadcast per owner into its slice of X_full (all_gather needs equal shapes)
        rank = dist.get_rank(group)
        full[int(bounds[rank]):int(bounds[rank + 1])].copy_(x_shard)
        for p in range(len(sizes)):
            if sizes[p]:
                dist.broadcast(full[int(bounds[p]):int(bounds[p + 1])], src=dist.get_global_rank(group, p) if group else p,
                               group=group)
    return full
