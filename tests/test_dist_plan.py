"""CPU tier: the index bookkeeping of the peer-memory multi-GPU path (csrc/dist.cu, restated in numpy as
gnnagg.partition.peer_plan), checked against the oracle: stage by stage accumulation over [own shard | receive
buffer] coordinates must reproduce the un-partitioned aggregation for every rank, world size and stage count."""
import numpy as np
import pytest

from gnnagg import partition, synth


def _stage_csr(ptr, plan, val, s):
    """sub-CSR of stage s in CSR order (what source_slices_build_device produces from the stage keys)"""
    n = len(ptr) - 1
    row = np.repeat(np.arange(n), np.diff(ptr))
    sel = plan["stage"] == s
    cnt = np.bincount(row[sel], minlength=n)
    p = np.zeros(n + 1, np.int32)
    p[1:] = np.cumsum(cnt)
    return p, np.ascontiguousarray(plan["idx_new"][sel]), np.ascontiguousarray(val[sel])


@pytest.mark.parametrize("world,stages", [(1, 1), (2, 1), (3, 1), (3, 2), (4, 3), (8, 2), (8, 7)])
def test_staged_plan_reproduces_the_full_aggregation(orc, world, stages):
    rng = np.random.default_rng(world * 10 + stages)
    sizes = rng.integers(20, 60, world)
    sizes[rng.integers(world)] = 0 if world > 2 else sizes[0]          # an empty shard is legal
    bounds = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n, F = int(bounds[-1]), 8
    ptr, idx = synth.small_random_csr(n, 7.0, 11 + world, hub=150)
    val = rng.standard_normal(len(idx)).astype(np.float32)
    X = rng.standard_normal((n, F)).astype(np.float32)
    want, scale = orc.spmm_f64(ptr, idx, val, X)
    for rank in range(world):
        lp, li, lv = partition.local_block(ptr, idx, val, bounds, rank)
        rows = len(lp) - 1
        plan = partition.peer_plan(li, bounds, rank, stages)
        assert plan["num_stages"] == (1 if world == 1 else 1 + min(stages, world - 1))
        assert sorted(plan["pull_order"]) == [p for p in range(world) if p != rank]
        # the receive buffer: rows pulled from the owners' shards, owner by owner
        recv = np.zeros((len(plan["recv_rows"]), F), np.float32)
        for p in range(world):
            a, b = plan["recv_off"][p], plan["recv_off"][p + 1]
            shard = X[bounds[p]:bounds[p + 1]]
            recv[a:b] = shard[plan["recv_local"][a:b]]
            assert p != rank or a == b                                  # own rows never travel
        assert np.array_equal(recv, X[plan["recv_rows"]])
        # stages are monotone along the pull order, every stage non-empty in owners
        st = [plan["stage_of"][p] for p in plan["pull_order"]]
        assert st == sorted(st) and (world == 1 or set(st) == set(range(1, plan["num_stages"])))
        acc = np.zeros((rows, F), np.float64)
        own = np.ascontiguousarray(X[bounds[rank]:bounds[rank + 1]])
        for s in range(plan["num_stages"]):
            p_s, i_s, v_s = _stage_csr(lp, plan, lv, s)
            src = own if s == 0 else recv
            if len(i_s):
                assert i_s.max() < len(src)
                y, _ = orc.spmm_f64(p_s, i_s, v_s, np.ascontiguousarray(src) if len(src) else np.zeros((1, F), np.float32))
                acc += y
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        assert np.all(np.abs(acc - want[lo:hi]) <= 1e-5 * scale[lo:hi] + 1e-30)


def test_every_rank_starts_with_a_different_owner():
    world = 8
    firsts = [partition.peer_plan(np.zeros(1, np.int32), np.arange(world + 1) * 4, r, 3)["pull_order"][0] for r in range(world)]
    assert sorted(firsts) == list(range(world))
