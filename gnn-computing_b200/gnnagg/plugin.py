"""plugin.py -- the reference's PyTorch extension surface (Figure7/kernel.cpp:37-179, module
`gnncompile` in Figure7/our.py:15-22) on top of the C ABI.

Same function names, argument order and meaning; handles are plain ints like the reference's
`int64_t` (there: a heap `Aggregator*`, never destroyed; here: a key into a registry, `destroy(at)` is
an addition).  `blocksize` is accepted and ignored.  Launches go to torch's CURRENT stream (the
reference uses the legacy default stream, which is torch's default current stream).

    import gnnagg.plugin as gnc
    ptrs, idxs = gnc.new_load(dset, "_thres_0.2", 0)
    at = gnc.gcn_init(ptrs, idxs, vals); gnc.gcn_schedule(at, 32)
    gnc.gcn_run(at, feat2, output_feat, 128, 1)
"""
import itertools

import torch

from . import Aggregator, GnnaggError, SCHED_NEIGHBOR_GROUPING, load_graph

_registry = {}
_ids = itertools.count(1)
n = -1  # the reference publishes the loaded sizes as globals (kernel.cpp:52-53)
m = -1
datadir = "../data/"  # load_graph hard-codes this relative path (src/data.cu:34)


def _get(at):
    try:
        return _registry[int(at)]
    except KeyError:
        raise GnnaggError("unknown aggregator handle %r" % (at,))


def _check(*tensors):
    for t in tensors:  # CHECK_INPUT of kernel.cpp:71-75
        if not t.is_cuda:
            raise GnnaggError("tensor must be a CUDA tensor")
        if not t.is_contiguous():
            raise GnnaggError("tensor must be contiguous")


def new_load(dset, reorder="", devid=0):
    """kernel.cpp:37-67: load_graph on the host (with optional .reorder<suffix>), two int32 CUDA tensors"""
    global n, m
    ptr, idx, _, _ = load_graph(dset, datadir, reorder)
    n, m = len(ptr) - 1, len(idx)
    dev = torch.device("cuda", devid)
    return [torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev)]


def gcn_init(ptrs, idxs, val):
    _check(ptrs, idxs, val)
    at = next(_ids)
    _registry[at] = Aggregator(ptrs, idxs, val)
    return at


def gcn_update_val(at, val):
    _check(val)
    _get(at).set_val(val)


def gcn_run(at, feat, outfeat, blocksize, scheduled):
    _check(feat, outfeat)
    _get(at).gcn_run(feat, outfeat, bool(scheduled))  # feature width = feat.size(1), kernel.cpp:100


def gcn_schedule(at, neighbor_num):
    _get(at).schedule(SCHED_NEIGHBOR_GROUPING, [int(neighbor_num)])


def gat_init(ptrs, idxs):
    _check(ptrs, idxs)
    at = next(_ids)
    _registry[at] = Aggregator(ptrs, idxs)
    return at


def gat_run(at, feat, att, outfeat, blocksize, scheduled):
    _check(feat, att, outfeat)
    _get(at).gat_run(feat, att, outfeat, 0.2, bool(scheduled))


def gat_run_u_add_v(at, att, outval, blocksize):
    _check(att, outval)
    _get(at).u_add_v(att, outval)


def gat_run_add_to_center(at, inval, outatt, blocksize):
    _check(inval, outatt)
    _get(at).add_to_center(inval, outatt)


def gat_run_div_each(at, inatt, inoutval, blocksize):
    _check(inatt, inoutval)
    _get(at).each_div(inatt, inoutval)


def gat_schedule(at, neighbor_num):
    _get(at).schedule(SCHED_NEIGHBOR_GROUPING, [int(neighbor_num)])


def destroy(at):
    _registry.pop(int(at)).close()
