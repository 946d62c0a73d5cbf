#!/usr/bin/env python
"""bench.py -- headline benchmark of the neighbour-aggregation hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

N = 1 (default): BASELINE.json configs[1] -- the fused GCN layer H = (A X) W, feat 128 -> 128, on the
synthetic reddit-shaped power-law graph (232,965 vertices / 114,615,891 edges), one B200.
N > 1 (launched by torch.distributed.run, one rank per GPU): the same layer on a graph that is
1-D row-partitioned by destination: every rank owns a reddit-shaped row block (232,965 rows,
114,615,891 edges whose sources span all N*232,965 vertices) and the matching X shard; a step is
the NCCL all-gather of the source-feature halo over NVLink followed by the local layer (weak scaling).

A "step" = one pass of the layer over the whole graph.  Reported metric: algorithmic GB/s of the
layer (gather model of SURVEY.md 8(d): 4(n+1) + 8m + 4mF_in + 4nF_out + 4F_inF_out bytes per rank).
`value`: inputs resident in HBM; `e2e`: the same step through the host-buffer entry point
(gnnagg_gcn_layer_host at N=1) with pinned HOST X/W/H and the copies inside the timed region.
`--impl reference`: the reference has no CPU implementation of this path; the arm times
the scalar CPU port (oracle/) on all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-computing_b200"))

WORKLOADS = {
    # name: (shape name, feat_in, feat_out)
    "reddit_gcn_layer_128": ("reddit", 128, 128),
    "arxiv_gcn_layer_32": ("arxiv", 32, 32),
    "proteins_gcn_layer_64": ("proteins", 64, 64),
    "products_gcn_layer_256": ("products", 256, 256),
    # BASELINE.json configs[4]: GCN aggregation only on the RMAT scale-26 graph (2^26 vertices / 2^30 edges), the
    # whole graph 1-D row-partitioned over the N ranks (strong scaling); a step = halo all-gather + aggregation
    "rmat26_gcn_agg_64": ("rmat26", 64, None),
    "rmat22_gcn_agg_64": ("rmat22", 64, None),   # down-scaled variant of the same workload
}
RMAT_SCALES = {"rmat26": (1 << 26, 1 << 30), "rmat22": (1 << 22, 1 << 26)}


def layer_bytes(n, m, fin, fout):
    """algorithmic bytes of one fused layer (or, fout=None, of the aggregation alone) on one rank
    (gather model)"""
    if fout is None:
        return spmm_bytes(n, m, fin)
    return 4 * (n + 1) + 4 * m + 4 * m + 4 * m * fin + 4 * n * fout + 4 * fin * fout


def spmm_bytes(n, m, F):
    return 4 * (n + 1) + 8 * m + 4 * m * F + 4 * n * F


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.active = False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
            while not self._stop_evt.is_set():
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    self.reasons.update(k for k, bit in names.items() if r & bit)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it rather than fake numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        self._stop_evt.set()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


def cpu_port_time(orc, hp, hi, hv, hX, hW, rows, fin, fout):
    """seconds of the scalar CPU port (oracle.c, all OpenMP threads) on rows [0, rows)"""
    import numpy as np

    AX = np.zeros((rows, fin), np.float32)
    H = np.zeros((rows, fout), np.float32) if fout is not None else None
    t0 = time.perf_counter()
    orc.spmm_f32(hp, hi, hv, hX, 0, rows, out=AX)
    if fout is not None:
        orc.dense_f32(AX, hW, 0, rows, out=H)
    return time.perf_counter() - t0


def cpu_sample(orc, ptr, idx, val, X, W, fin, fout, budget_s):
    """times the CPU port on a leading row block sized for ~budget_s seconds; returns
    (GB/s of the sample, seconds, rows, edges)"""
    import numpy as np

    n = ptr.numel() - 1
    hp_full = ptr.cpu().numpy()
    hX, hW = X.cpu().numpy(), W.cpu().numpy()

    def block(rows):
        e = int(hp_full[rows])
        return np.ascontiguousarray(hp_full[: rows + 1]), idx[:e].cpu().numpy(), val[:e].cpu().numpy(), e

    probe_rows = int(np.searchsorted(hp_full, min(int(hp_full[-1]), 2_000_000), side="left"))
    probe_rows = max(1, min(n, probe_rows))
    hp, hi, hv, e = block(probe_rows)
    cpu_port_time(orc, hp, hi, hv, hX, hW, probe_rows, fin, fout)  # warm-up (page faults, OpenMP pool)
    t = cpu_port_time(orc, hp, hi, hv, hX, hW, probe_rows, fin, fout)
    rate = max(e, 1) / max(t, 1e-6)  # edges/s
    want_edges = int(min(int(hp_full[-1]), rate * budget_s))
    rows = max(probe_rows, min(n, int(np.searchsorted(hp_full, want_edges, side="left"))))
    hp, hi, hv, e = block(rows)
    t = min(cpu_port_time(orc, hp, hi, hv, hX, hW, rows, fin, fout) for _ in range(2))
    return layer_bytes(rows, e, fin, fout) / t / 1e9, t, rows, e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit_gcn_layer_128", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--scheduled", type=int, default=0, help="1: run the neighbour-grouped (NG=32) path")
    ap.add_argument("--sources", default="rmat", choices=["rmat", "uniform"],
                    help="rmat: the R-MAT source distribution of the workload definition (default, headline); uniform: same "
                         "degree sequence but uniformly random sources -- the cache-hostile extreme, reported as context")
    ap.add_argument("--halo", default="auto", choices=["auto", "allgather", "pruned"],
                    help="N>1: all-gather the whole X, or exchange only the referenced source rows (all-to-all). "
                         "auto = pruned for the partitioned R-MAT graph, all-gather otherwise")
    ap.add_argument("--pipeline", type=int, default=0,
                    help="N>1: number of row chunks of the pipelined halo all-gather (0 = one all-gather, then the layer)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    shape, fin, fout = WORKLOADS[args.workload]

    from gnnagg import synth

    strong = shape in RMAT_SCALES
    if strong:  # one fixed graph, partitioned: per-rank block = total / N
        n_tot, m_tot = RMAT_SCALES[shape]
        n, m = n_tot // args.gpus, m_tot // args.gpus
        what = "GCN aggregation (CSR SpMM) feat %d on the R-MAT %s graph (%d vertices / %d edges in total), %d rows / %d edges per GPU" % (
            fin, shape, n_tot, m_tot, n, m)
    else:
        n, m = synth.shape_of(shape)
        what = "fused GCN layer (CSR SpMM aggregation + dense combination) feat %d->%d on synthetic %s-shaped R-MAT graph, " \
               "%d vertices / %d edges per GPU" % (fin, fout, shape, n, m)
    config = {"workload": "%s: %s" % (args.workload, what),
              "graph": "rmat(a=.57,b=.19,c=.19,d=.05) seed=123, val=1/sqrt((deg_u+1)(deg_v+1)), X~N(0,1), W~N(0,1)/sqrt(F)",
              "partition": "1d-row-by-destination" if args.gpus > 1 else "single-gpu",
              "scheduled": bool(args.scheduled), "sources": args.sources,
              "l2": "flushed between timed steps (256 MiB write) and inputs (idx+val %.0f MB, X %.0f MB) exceed the 126 MB L2"
                    % (8 * m / 1e6, 4 * n * fin / 1e6)}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import oracle as orc

        orc.use_all_cores()  # rank 0 alone runs this arm: every host core, also under torchrun (OMP_NUM_THREADS=1 there)

        dev = torch.device("cuda:%d" % local_rank) if torch.cuda.is_available() else torch.device("cpu")
        # the same graph; a leading row block is what gets timed, so only that block is built when no GPU is around
        gen_rows, gen_edges = (n, m) if dev.type == "cuda" else (n, min(m, 4_000_000))
        src_total = n * args.gpus if strong else n  # the reference arm always times rank 0's block
        ptr, idx = synth.rmat_csr(gen_rows, gen_edges, seed=123, device=dev, src_num_v=src_total if strong and args.gpus > 1 else None,
                                  dst_prefix=0 if strong and args.gpus > 1 else None)
        val = synth.gcn_norm_val(ptr, idx) if src_total == n else torch.rand(idx.numel(), device=dev) + 0.5
        g = torch.Generator(device=dev).manual_seed(123)
        X = torch.randn((src_total, fin), device=dev, generator=g)
        W = torch.randn((fin, fout or fin), device=dev, generator=g) / fin ** 0.5
        per_step = []
        info = None
        for it in range(args.warmup + args.steps):
            budget = max(1.0, min(args.cpu_seconds, 120.0 / (args.warmup + args.steps)))
            gbs, t, rows, e = cpu_sample(orc, ptr, idx, val, X, W, fin, fout, budget) if it == 0 else info
            if it == 0:
                info = (gbs, t, rows, e)
                hp = ptr[: rows + 1].cpu().numpy()
                hi, hv = idx[:e].cpu().numpy(), val[:e].cpu().numpy()
                hX, hW = X.cpu().numpy(), W.cpu().numpy()
            t = cpu_port_time(orc, hp, hi, hv, hX, hW, rows, fin, fout)
            if it >= args.warmup:
                per_step.append(t)
        t_mean = float(np.mean(per_step))
        gbs = layer_bytes(rows, e, fin, fout) / t_mean / 1e9
        sample = "rows [0,%d) of the workload graph = %d edges (%.2f%% of m), scalar fp32 CSR port + fp32 GEMM, OpenMP" % (
            rows, e, 100.0 * e / m)
        line = {"impl": "reference", "metric": "gcn_layer_algorithmic_GBps", "value": round(gbs, 3), "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_mean * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": orc.num_threads(), "kind": "port",
                                 "sample": sample},
                "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "the reference has no CPU implementation of the float path (SURVEY 8(c)); this is the CPU port "
                        "in oracle/oracle.c; each step processes the bounded sample, GB/s is per sample bytes"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU port"
    import gnnagg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda:%d" % local_rank)
    dist = None
    if args.gpus > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    N = args.gpus
    src_n = n * N

    # graph: rank r owns rows [r*n, (r+1)*n) of an (N*n)-vertex R-MAT graph, sources are global ids
    t0 = time.time()
    ptr, idx = synth.rmat_csr(n, m, seed=123, device=dev, src_num_v=src_n if N > 1 else None,
                              dst_prefix=rank if N > 1 else None)
    if args.sources == "uniform":
        idx = torch.randint(0, src_n, (m,), device=dev, dtype=torch.int32,
                            generator=torch.Generator(device=dev).manual_seed(99 + rank))
    deg = (ptr[1:] - ptr[:-1]).to(torch.float32)
    if N > 1:
        deg_all = torch.empty(src_n, device=dev)
        dist.all_gather_into_tensor(deg_all, deg)
        val = synth.gcn_norm_val(ptr, idx, src_deg=deg_all)
        del deg_all
    else:
        val = synth.gcn_norm_val(ptr, idx)
    g = torch.Generator(device=dev).manual_seed(123 + rank)
    Xs = torch.randn((n, fin), device=dev, generator=g)          # this rank's X shard
    agg_only = fout is None
    W = torch.randn((fin, fout or fin), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / fin ** 0.5
    halo = args.halo
    if halo == "auto" and args.gpus > 1:
        # exchange only the referenced source rows when that is clearly less than X (decided from the data, identically
        # on every rank: the fraction of rank 0 is broadcast)
        import torch.distributed as dist0

        frac = torch.tensor([torch.unique(idx).numel() / float(src_n)], device=dev)
        dist0.broadcast(frac, src=0)
        halo = "pruned" if float(frac.item()) < 0.7 else "allgather"
    elif halo == "auto":
        halo = "allgather"
    pruned = N > 1 and halo == "pruned" and not args.scheduled
    pipelined = N > 1 and args.pipeline > 0 and not args.scheduled and not pruned
    Xfull = torch.empty((src_n, fin), device=dev) if (N > 1 and not pipelined and not pruned) else Xs
    H = torch.empty((n, fout or fin), device=dev)
    ph = None
    if pruned:
        from gnnagg.partition import PrunedHalo

        if args.pipeline > 0:
            from gnnagg.partition import PipelinedPrunedHalo

            ph = PipelinedPrunedHalo(ptr, idx, val, n, N, rank, fin, chunks=args.pipeline)
            agg = ph.aggs[0]
            config["halo"] = "pruned + pipelined: %d row chunks, chunk c fetches only the referenced source rows no earlier chunk " \
                             "fetched (all-to-all per chunk, %.1f%% of X on rank 0, stage shares %s) while chunk c-1 is aggregated" % (
                                 args.pipeline, 100 * ph.referenced_fraction, [round(f, 2) for f in ph.stage_fraction])
        else:
            ph = PrunedHalo(ptr, idx, val, n, N, rank, fin)
            agg = ph.agg
            config["halo"] = "pruned: only referenced source rows travel (all-to-all, %.1f%% of X on rank 0), CSR re-indexed " \
                             "into the compact receive buffer" % (100 * ph.referenced_fraction)
    else:
        agg = gnnagg.Aggregator(ptr, idx, val)
    if args.scheduled:
        agg.schedule(gnnagg.SCHED_NEIGHBOR_GROUPING, [32])
    pipe = AX = None
    AXp = torch.empty((n, fin), device=dev) if (pruned and args.pipeline > 0 and not agg_only) else None
    if pipelined:
        from gnnagg.partition import HaloPipeline

        pipe = HaloPipeline(ptr, idx, val, n, N, rank, fin, chunks=args.pipeline)
        AX = torch.empty((n, fin), device=dev)
        config["halo"] = "all-gather cut into %d row chunks, sub-CSR of chunk c accumulated while chunk c+1 is in flight; " \
                         "edges with local sources first" % args.pipeline
    elif N > 1 and not pruned:
        config["halo"] = "one NCCL all-gather of X, then the layer"
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def layer(x_full):
        if pipelined:
            pipe.aggregate(Xs, H if agg_only else AX)  # chunked all-gather overlapped with aggregation
            if not agg_only:
                gnnagg.dense_nn(AX, W, H)
        else:
            if pruned and args.pipeline > 0:
                ph.aggregate(Xs, H if agg_only else AXp)   # pack, per-chunk all-to-all overlapped with aggregation
                if not agg_only:
                    gnnagg.dense_nn(AXp, W, H)
                return
            if pruned:
                x_full = ph.exchange(Xs)                  # pack + all-to-all of the referenced rows
            elif N > 1:
                dist.all_gather_into_tensor(Xfull, Xs)   # source-feature halo over NVLink
            if agg_only:
                agg.gcn_run(x_full, H, scheduled=bool(args.scheduled))
            else:
                agg.gcn_layer(x_full, W, H, None, scheduled=bool(args.scheduled))

    def step():
        layer(Xfull)

    # host buffers for the end-to-end number
    hX = torch.empty((n, fin), pin_memory=True).copy_(Xs)
    hW = torch.empty((fin, fout or fin), pin_memory=True).copy_(W)
    hH = torch.empty((n, fout or fin), pin_memory=True)

    def step_e2e():
        if N == 1 and agg_only:
            agg.gcn_run_host(hX, hH, scheduled=bool(args.scheduled))
        elif N == 1:
            agg.gcn_layer_host(hX, hW, hH, scheduled=bool(args.scheduled))  # H2D + layer + D2H + sync inside
        else:
            Xs.copy_(hX, non_blocking=True)
            layer(Xfull)
            hH.copy_(H, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        prof = []
        aggs = pipe.aggs if pipelined else (ph.aggs if (pruned and args.pipeline > 0) else [agg])
        count = lambda: sum(a.launches for a in aggs)
        l0 = count()
        sampler.active = True
        for a, b in evs:
            flush.fill_(1)  # L2 flush, outside the timed events
            a.record()
            fn()
            b.record()
            if profile:
                reads = [a.profile_read() for a in aggs]
                prof.append({k: sum(r[k] for r in reads) for k in reads[0]})
        barrier()
        sampler.active = False
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if dist is not None:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps, count() - l0 + (steps if (pipelined and not agg_only) else 0), prof  # + the dense launch

    sampler = ClockSampler(local_rank)
    sampler.start()
    prof_aggs = pipe.aggs if pipelined else (ph.aggs if (pruned and args.pipeline > 0) else [agg])
    for a_ in prof_aggs:
        a_.profile(True)
    ms, launches, prof = timed(step, args.steps, args.warmup, profile=True)
    for a_ in prof_aggs:
        a_.profile(False)
    ms_e2e, _, _ = timed(step_e2e, max(3, args.steps // 2), 3)
    clocks = sampler.result()

    bytes_rank = layer_bytes(n, m, fin, fout)
    value = N * bytes_rank / (ms * 1e-3) / 1e9
    e2e_value = N * bytes_rank / (ms_e2e * 1e-3) / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    agg_ms = float(np.mean([p["agg"] for p in prof]))
    traffic, ncu_units = None, {}
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload
        summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        key = args.workload + ("_sched" if args.scheduled else "") + ("_uniform" if args.sources == "uniform" else "")
        traffic = summ.get(key, {}).get("dram_bytes_per_launch")
        ncu_units = {k: summ[key][k] for k in ("l1tex_throughput_pct", "lts_throughput_pct", "dram_GBps", "l2_hit_pct",
                                               "l1_hit_pct", "duration_ms", "source") if k in summ.get(key, {})}
    except Exception:
        pass
    achieved = spmm_bytes(n, m, fin) / (agg_ms * 1e-3) / 1e9
    lpr, nv = (8, 1) if fin <= 32 else (16, 1) if fin <= 64 else (32, 1) if fin <= 128 else (32, 2)
    roofline = {"bound": "hbm", "kernel": "agg_kernel<%d,%d,GCN,%s,%d> (CSR SpMM aggregation)" % (
                    lpr, nv, "sched" if args.scheduled else "csr", 128 if m < 4000000 else 512),
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src,
                # which unit binds, from the committed `ncu --set full` capture of this kernel on this workload (percent of
                # the unit's peak throughput; DRAM in GB/s): the gathers run out of L1/L2, so HBM is not the binding unit here
                "ncu": ncu_units,
                "algorithmic_bytes_per_launch": spmm_bytes(n, m, fin), "kernel_ms": round(agg_ms, 4),
                "step_breakdown_ms": {k: round(float(np.mean([p[k] for p in prof])), 4) for k in ("agg", "agg_rest", "dense", "total")},
                "note": "gather model: one F-float source row per edge; X (%.0f MB) is L2-resident to a large degree, so this is an "
                        "effective-bandwidth figure and may exceed the HBM copy peak" % (4 * n * fin / 1e6)}

    import oracle as orc

    orc.use_all_cores()
    if N > 1:  # the CPU port needs the replicated X of rank 0's block
        g_all = [torch.randn((n, fin), device=dev, generator=torch.Generator(device=dev).manual_seed(123 + r)) for r in range(N)]
        Xcpu = torch.cat(g_all)
    else:
        Xcpu = Xs
    gbs, t_cpu, rows, e = cpu_sample(orc, ptr, idx, val, Xcpu, W, fin, fout, args.cpu_seconds)
    cpu_baseline = {"value": round(gbs, 3), "unit": "GB/s", "cores": orc.num_threads(), "kind": "port",
                    "sample": "rows [0,%d) = %d edges (%.2f%% of m) of rank 0's graph, %.1f s; scalar fp32 CSR port + fp32 GEMM "
                              "(oracle/oracle.c, OpenMP dynamic,64)" % (rows, e, 100.0 * e / m, t_cpu)}

    line = {"metric": "gcn_aggregation_algorithmic_GBps" if agg_only else "gcn_layer_algorithmic_GBps", "value": round(value, 1),
            "unit": "GB/s", "n_gpus": N,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "edges_feat_per_s": round(N * m * fin / (ms * 1e-3), 1),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "GB/s", "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": 4 * n * fin + (4 * fin * fout if (N == 1 and not agg_only) else 0),
                    "d2h_bytes_per_step": 4 * n * (fout or fin),
                    "api": ("gnnagg_gcn_run_host (pinned host X -> Y)" if agg_only else "gnnagg_gcn_layer_host (pinned host X, W -> H)") if N == 1 else
                           "pinned H2D of the X shard + NCCL halo all-gather + aggregation + combination + D2H of the H shard"},
            "gpu_launches": int(launches) + (args.steps if pruned else 0),  # + the row-packing kernel
            "compute_only": {"ms_per_step": round(float(np.mean([p["total"] for p in prof])), 4),
                             "value": round(N * bytes_rank / (float(np.mean([p["total"] for p in prof])) * 1e-3) / 1e9, 1), "unit": "GB/s",
                             "note": "rank 0's kernels only (X already resident, no halo exchange), from the library's own CUDA events"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "setup_s": round(t_setup, 2)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
