"""ctypes front-end of the CPU oracle (oracle/oracle.c) and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(ref=True):
    """compile oracle.c (and the reference, when /root/reference exists)"""
    subprocess.check_call(["make", "-s", "-C", _HERE] + ([] if ref else [os.path.join(_HERE, "_build", "liboracle.so")]))


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build(ref=False)
        _lib = C.CDLL(_LIB)
        _lib.orc_neighbor_grouping.restype = C.c_int64
        _lib.orc_locality.restype = C.c_int64
    return _lib


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)"""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(cores))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ float path
def spmm_f32(ptr, idx, val, X, row_begin=0, row_end=None, out=None):
    n = len(ptr) - 1
    F = X.shape[1]
    row_end = n if row_end is None else row_end
    Y = np.zeros((n, F), np.float32) if out is None else out
    lib().orc_spmm_f32(C.c_int64(row_begin), C.c_int64(row_end), _vp(ptr), _vp(idx), _vp(val), _vp(X), C.c_int(F), _vp(Y))
    return Y


def spmm_f64(ptr, idx, val, X):
    n = len(ptr) - 1
    F = X.shape[1]
    Y = np.empty((n, F), np.float32)
    S = np.empty((n, F), np.float32)
    lib().orc_spmm_f64(C.c_int64(n), _vp(ptr), _vp(idx), _vp(val), _vp(X), C.c_int(F), _vp(Y), _vp(S))
    return Y, S


def spmm_grouped_f64(n, gptr, gidx, gval, target, X):
    F = X.shape[1]
    Y = np.empty((n, F), np.float32)
    S = np.empty((n, F), np.float32)
    lib().orc_spmm_grouped_f64(C.c_int64(n), C.c_int64(len(target)), _vp(gptr), _vp(gidx), _vp(gval), _vp(target),
                               _vp(X), C.c_int(F), _vp(Y), _vp(S))
    return Y, S


def csr2edgelist(ptr, idx):
    out = np.empty(2 * len(idx), np.int32)
    lib().orc_csr2edgelist(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(out))
    return out


def dense_f64(A, W):
    M, K = A.shape
    N = W.shape[1]
    H = np.empty((M, N), np.float32)
    S = np.empty((M, N), np.float32)
    lib().orc_dense_f64(C.c_int64(M), C.c_int(K), C.c_int(N), _vp(A), _vp(W), _vp(H), _vp(S))
    return H, S


def dense_f32(A, W, row_begin=0, row_end=None, out=None):
    M, K = A.shape
    N = W.shape[1]
    row_end = M if row_end is None else row_end
    H = np.zeros((M, N), np.float32) if out is None else out
    lib().orc_dense_f32(C.c_int64(row_begin), C.c_int64(row_end), C.c_int(K), C.c_int(N), _vp(A), _vp(W), _vp(H))
    return H


def gcn_layer_f64(ptr, idx, val, X, W):
    n = len(ptr) - 1
    K = X.shape[1]
    N = W.shape[1]
    AX = np.empty((n, K), np.float32)
    H = np.empty((n, N), np.float32)
    S = np.empty((n, N), np.float32)
    lib().orc_gcn_layer_f64(C.c_int64(n), _vp(ptr), _vp(idx), _vp(val), _vp(X), C.c_int(K), _vp(W), C.c_int(N),
                            _vp(AX), _vp(H), _vp(S))
    return AX, H, S


def gat_f64(ptr, idx, att, X, slope=0.2, empty_value=0.0):
    n = len(ptr) - 1
    F = X.shape[1]
    Y = np.empty((n, F), np.float32)
    den = np.empty(n, np.float32)
    S = np.empty((n, F), np.float32)
    lib().orc_gat_f64(C.c_int64(n), _vp(ptr), _vp(idx), _vp(att), C.c_float(slope), _vp(X), C.c_int(F),
                      C.c_float(empty_value), _vp(Y), _vp(den), _vp(S))
    return Y, den, S


def edge_softmax_f64(ptr, idx, att, slope=0.2):
    out = np.empty(len(idx), np.float32)
    lib().orc_edge_softmax_f64(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(att), C.c_float(slope), _vp(out))
    return out


def edge_weight_f64(ptr, idx, att, slope=0.2):
    out = np.empty(len(idx), np.float32)
    lib().orc_edge_weight_f64(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(att), C.c_float(slope), _vp(out))
    return out


def u_add_v(ptr, idx, att):
    out = np.empty(len(idx), np.float32)
    lib().orc_u_add_v(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(att), _vp(out))
    return out


def add_to_center_f64(ptr, newval):
    out = np.empty(len(ptr) - 1, np.float32)
    lib().orc_add_to_center_f64(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(newval), _vp(out))
    return out


def each_div(ptr, center, newval):
    out = newval.copy()
    lib().orc_each_div(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(center), _vp(out))
    return out


def mlp_f64(ptr, idx, X, W):
    n = len(ptr) - 1
    F = X.shape[1]
    Y = np.empty((n, F), np.float32)
    S = np.empty((n, F), np.float32)
    lib().orc_mlp_f64(C.c_int64(n), _vp(ptr), _vp(idx), _vp(X), _vp(W), C.c_int(F), _vp(Y), _vp(S))
    return Y, S


def sddmm_f64(ptr, idx, X1, X2):
    F = X1.shape[1]
    out = np.empty(len(idx), np.float32)
    S = np.empty(len(idx), np.float32)
    lib().orc_sddmm_f64(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(X1), _vp(X2), C.c_int(F), _vp(out), _vp(S))
    return out, S


# ------------------------------------------------------------------ samplers
def sample_subgraph(ptr, idx, active, fanout=0, layer_num=1, seed=123):
    """(active_after, vertexset, sub_ptr, sub_idx); fanout <= 0 restates sampleVertex (sample.h:131-200)"""
    n = len(ptr) - 1
    act = np.ascontiguousarray(active, np.int32).copy()
    vs, sp, si = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(len(idx), 1), np.int32)
    ne = C.c_int()
    rows = lib().orc_sample_subgraph(C.c_int(n), _vp(ptr), _vp(idx), _vp(act), C.c_int(fanout), C.c_int(layer_num),
                                     C.c_uint64(seed), _vp(vs), _vp(sp), _vp(si), C.c_int64(len(si)), C.byref(ne))
    assert rows >= 0
    return act, vs[:rows].copy(), sp[: rows + 1].copy(), si[: ne.value].copy()


def sample_pos(seed, v, j, deg, fanout):
    return lib().orc_sample_pos(C.c_uint64(seed), C.c_int(v), C.c_int(j), C.c_int(deg), C.c_int(fanout))


# ------------------------------------------------------------------ backward
def transpose_csr(ptr, idx, num_src):
    m = len(idx)
    t_ptr, t_idx, t_perm = np.empty(num_src + 1, np.int32), np.empty(m, np.int32), np.empty(m, np.int32)
    lib().orc_transpose_csr(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), C.c_int(num_src), _vp(t_ptr), _vp(t_idx), _vp(t_perm))
    return t_ptr, t_idx, t_perm


def spmm_t_f64(ptr, idx, val, dY, num_src):
    F = dY.shape[1]
    dX, S = np.empty((num_src, F), np.float32), np.empty((num_src, F), np.float32)
    lib().orc_spmm_t_f64(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(val), _vp(dY), C.c_int(F), C.c_int(num_src),
                         _vp(dX), _vp(S))
    return dX, S


def gat_backward_f64(ptr, idx, att, X, dY, slope=0.2):
    """(dX, datt, dX_scale, datt_scale); att / datt have X.shape[0] = max(n, num_src) rows"""
    n, (num_src, F) = len(ptr) - 1, X.shape
    rows = att.shape[0]
    assert rows >= n and rows >= num_src
    dX, SX = np.empty((num_src, F), np.float32), np.empty((num_src, F), np.float32)
    dA, SA = np.empty((rows, 2), np.float32), np.empty((rows, 2), np.float32)
    lib().orc_gat_backward_f64(C.c_int64(n), _vp(ptr), _vp(idx), _vp(att), C.c_float(slope), _vp(X), _vp(dY), C.c_int(F),
                               C.c_int(num_src), C.c_int64(rows), _vp(dX), _vp(dA), _vp(SX), _vp(SA))
    return dX, dA, SX, SA


def gat_loss_f64(ptr, idx, att64, X64, dY64, slope=0.2):
    """sum(dY * Y(att, X)) with fp64 inputs: the scalar the finite-difference test differentiates"""
    f = lib().orc_gat_loss_f64
    f.restype = C.c_double
    return f(C.c_int64(len(ptr) - 1), _vp(ptr), _vp(idx), _vp(att64), C.c_double(slope), _vp(X64), _vp(dY64),
             C.c_int(X64.shape[1]))


# ------------------------------------------------------------------ integer path
def neighbor_grouping(ptr, idx, neighbor_num):
    n, m = len(ptr) - 1, len(idx)
    L = lib()
    g = L.orc_neighbor_grouping(_vp(ptr), _vp(idx), neighbor_num, n, m, None, None, None)
    op, oi, ot = np.empty(g + 1, np.int32), np.empty(m, np.int32), np.empty(g, np.int32)
    L.orc_neighbor_grouping(_vp(ptr), _vp(idx), neighbor_num, n, m, _vp(op), _vp(oi), _vp(ot))
    return op, oi, ot


def locality(ptr, idx, par_num, total_num_v, val=None, neighbor_num=0):
    """neighbor_num == 0: locality_schedule; > 0: localityNeighborGrouping"""
    n, m = len(ptr) - 1, len(idx)
    L = lib()
    g = L.orc_locality(_vp(ptr), _vp(idx), _opt(val), par_num, neighbor_num, n, total_num_v, None, None, None, None)
    op, oi, ot = np.empty(g + 1, np.int32), np.empty(m, np.int32), np.empty(g, np.int32)
    ov = np.empty(m, np.float32) if val is not None else None
    L.orc_locality(_vp(ptr), _vp(idx), _opt(val), par_num, neighbor_num, n, total_num_v, _vp(op), _vp(oi), _vp(ot),
                   _opt(ov))
    # sources outside [0,total_num_v) are dropped by the reference too: trim to what was emitted
    k = int(op[-1]) if g > 0 else 0
    return op, oi[:k], ot, (ov[:k] if ov is not None else None)


def reorder_csr(ptr, idx, rows, reverse_rows):
    newptr = np.empty_like(ptr)
    newidx = np.empty_like(idx)
    lib().orc_reorder_csr(_vp(ptr), _vp(idx), _vp(rows), _vp(reverse_rows), len(ptr) - 1, _vp(newptr), _vp(newidx))
    return newptr, newidx


def load_graph(datadir, dset, reorder_path=None):
    L = lib()
    nv, ne = C.c_int(), C.c_int()
    rc = L.orc_load_config(datadir.encode(), dset.encode(), C.byref(nv), C.byref(ne))
    if rc:
        raise IOError("config %s%s.config: rc=%d" % (datadir, dset, rc))
    ptr = np.empty(nv.value + 1, np.int32)
    idx = np.empty(ne.value, np.int32)
    rc = L.orc_load_graph(datadir.encode(), dset.encode(), nv.value, ne.value, _vp(ptr), _vp(idx))
    if rc:
        raise IOError("graph %s%s.graph: rc=%d" % (datadir, dset, rc))
    rows = rev = None
    if reorder_path and os.path.exists(reorder_path):
        rows = np.empty(nv.value, np.int32)
        rev = np.empty(nv.value, np.int32)
        rc = L.orc_read_reorder(reorder_path.encode(), nv.value, _vp(rows), _vp(rev))
        if rc:
            raise IOError("reorder %s: rc=%d" % (reorder_path, rc))
        ptr, idx = reorder_csr(ptr, idx, rows, rev)
    return ptr, idx, rows, rev


def validate2(ref, ans):
    with np.errstate(all="ignore"):
        return int(lib().orc_validate2(_vp(ref), _vp(ans), C.c_int64(ref.size)))


def validate_reordered(ref, ans, rows):
    return int(lib().orc_validate_reordered(_vp(ref), _vp(ans), _vp(rows), ref.shape[0], ref.shape[1]))


# ------------------------------------------------------------------ the compiled reference
_ref = None


def ref_available():
    return os.path.exists(_REF)


def ref():
    """oracle/_ref/libref.so: the reference's own sources compiled by oracle/Makefile"""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF):
            raise RuntimeError("oracle/_ref/libref.so missing (run `make -C oracle`; needs /root/reference)")
        _ref = C.CDLL(_REF)
        _ref.ref_sched_run.restype = C.c_longlong
        _ref.ref_sched_num_edges.restype = C.c_longlong
        _ref.ref_sched_num_ptr.restype = C.c_longlong
        for f in ("ref_gcn_create", "ref_gat_create", "ref_sddmm_create", "ref_mlp_create"):
            getattr(_ref, f).restype = C.c_void_p
        _ref.ref_gcn_run_edgewise.restype = C.c_double
        _ref.ref_sddmm_run.restype = C.c_double
        _ref.ref_mlp_run.restype = C.c_double
    return _ref


def ref_schedule(kind, ptr, idx, val=None, par_num=1, neighbor_num=1, total_num_v=None):
    """run the reference's own graph_schedule.h function; kind as enum Schedule (0,1,2)"""
    R = ref()
    n, m = len(ptr) - 1, len(idx)
    total = n if total_num_v is None else total_num_v
    g = R.ref_sched_run(kind, _vp(ptr), _vp(idx), _opt(val), par_num, neighbor_num, n, m, total)
    ne = R.ref_sched_num_edges()
    op, oi, ot = np.empty(R.ref_sched_num_ptr(), np.int32), np.empty(ne, np.int32), np.empty(g, np.int32)
    ov = np.empty(ne, np.float32) if (val is not None and kind != 1) else None
    R.ref_sched_fetch(_vp(op), _vp(oi), _vp(ot), _opt(ov))
    return op, oi, ot, ov


def ref_reorder_csr(ptr, idx, rows, reverse_rows):
    newptr = np.empty_like(ptr)
    newidx = np.empty_like(idx)
    ref().ref_reorder_csr(_vp(ptr), _vp(idx), _vp(rows), _vp(reverse_rows), len(ptr) - 1, len(idx), _vp(newptr),
                          _vp(newidx))
    return newptr, newidx


def ref_load_graph(dset, reorder_subfix=""):
    """reference load_graph; reads ../data/<dset>.* relative to the CWD"""
    R = ref()
    nv, ne = C.c_int(), C.c_int()
    R.ref_load_graph(dset.encode(), reorder_subfix.encode(), C.byref(nv), C.byref(ne))
    ptr, idx = np.empty(nv.value + 1, np.int32), np.empty(ne.value, np.int32)
    rows, rev = np.empty(nv.value, np.int32), np.empty(nv.value, np.int32)
    R.ref_load_fetch(_vp(ptr), _vp(idx), _vp(rows), _vp(rev))
    if not R.ref_load_was_reordered():
        rows = rev = None
    return ptr, idx, rows, rev
