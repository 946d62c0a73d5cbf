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


@pytest.mark.parametrize("world,stages", [(1, 1), (2, 0), (2, 1), (3, 1), (3, 2), (4, 0), (4, 3), (8, 2), (8, 7), (2, -3), (4, -4), (8, -16)])
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
        plan = partition.peer_plan(li, bounds, rank, stages, ptr=lp)
        assert plan["num_stages"] == (1 if world == 1 else 1 + max(0, min(stages, world - 1)))
        assert sorted(plan["recv_order"]) == [p for p in range(world) if p != rank]
        # the receive slots: rows pushed by the owners out of their shards, owner by owner
        recv = np.zeros((len(plan["recv_rows"]), F), np.float32)
        for c in range(plan["rounds"]):
            for p in range(world):
                a, b = plan["recv_tab"][c][p], plan["recv_tab"][c][p + 1]
                shard = X[bounds[p]:bounds[p + 1]]
                recv[a:b] = shard[plan["recv_local"][a:b]]
                assert p != rank or a == b                              # own rows never travel
        assert np.array_equal(recv, X[plan["recv_rows"]])
        if plan["rounds"] > 1:
            # row pipelining: chunk c only gathers rows that arrived in rounds <= c
            row = np.repeat(np.arange(rows), np.diff(lp))
            chunk_e = np.searchsorted(plan["chunk_rows"][1:plan["rounds"]], row, side="right")
            rem = plan["idx_new"] >= rows
            limit = plan["recv_tab"][:, world][chunk_e[rem]]            # end of the round of the edge's chunk
            assert np.all(plan["idx_new"][rem] - rows < limit)
            cr = plan["chunk_rows"]
            assert cr[0] == 0 and cr[-1] == rows and np.all(np.diff(cr) >= 0)
        # stages are monotone along the arrival order, every stage non-empty in owners
        st = [plan["stage_of"][p] for p in plan["recv_order"]]
        assert st == sorted(st) and (world == 1 or stages <= 0 or set(st) == set(range(1, plan["num_stages"])))
        acc = np.zeros((rows, F), np.float64)
        # ONE buffer serves every stage: the own shard, then the receive slots
        src = np.ascontiguousarray(np.concatenate([X[bounds[rank]:bounds[rank + 1]], recv])) if rows + len(recv) else np.zeros((1, F), np.float32)
        for s in range(plan["num_stages"]):
            p_s, i_s, v_s = _stage_csr(lp, plan, lv, s)
            if len(i_s):
                assert i_s.max() < len(src)
                if s == 0 and plan["num_stages"] > 1:
                    assert i_s.max() < rows                           # stage 0 only touches local rows
                y, _ = orc.spmm_f64(p_s, i_s, v_s, src)
                acc += y
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        assert np.all(np.abs(acc - want[lo:hi]) <= 1e-5 * scale[lo:hi] + 1e-30)


def test_every_receiver_is_written_by_one_owner_at_a_time():
    """owner p pushes to p-1, p-2, ...; receiver q expects q+1, q+2, ...: in slot k the pairs form a permutation"""
    world = 8
    orders = [partition.peer_plan(np.zeros(1, np.int32), np.arange(world + 1) * 4, r, 3)["recv_order"] for r in range(world)]
    for k in range(world - 1):
        owners = [orders[q][k] for q in range(world)]
        assert sorted(owners) == list(range(world))
        for q in range(world):
            p = owners[q]
            assert (p - 1 - k) % world == q       # what dist.cu's push loop computes on the owner side
