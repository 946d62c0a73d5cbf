// ref_wrap.cu -- flat extern "C" doorway onto the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  This translation unit is compiled
// together with /root/reference/src/util.cu and /root/reference/src/data.cu, and includes the
// reference headers from where they lie (-I/root/reference/include); the result goes to
// oracle/_ref/libref.so (git-ignored).  No reference source is copied into this repository.
//
// What it exposes:
//   * the reference's host schedule / reorder / loader functions (run on CPU, no GPU needed)
//     -> used to pin oracle.c and to generate tests/golden/*.json
//   * the reference's own CUDA aggregators recompiled for sm_100
//     -> second oracle for the float path on the GPU box and the "kernel to beat" in bench.py
#include "util.h"
#include "data.h"
#include "aggr_gcn.h"
#include "aggr_gat.h"
#include "aggr_sddmm.h"
#include "aggr_nn.h"
#include "dense.h"
#include "sample.h"

#include <cstring>
// defined in src/data.cu:4 but not declared by include/data.h (its prototype there has 4 parameters)
void reorderCSR(const int *ptr, const int *idx, const int *map, const int *reverse_map, int num_v, int num_e,
                int *&newptr, int *&newidx);
#include <vector>

namespace {
std::vector<int> g_ptr, g_idx, g_target;
std::vector<float> g_val;
void clear_sched()
{
    g_ptr.clear();
    g_idx.clear();
    g_target.clear();
    g_val.clear();
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- host: schedules
// kind: 0 locality, 1 neighbor_grouping, 2 locality_neighbor_grouping (enum Schedule order)
long long ref_sched_run(int kind, int *ptr, int *idx, float *val, int par_num, int neighbor_num, int num_v, int num_e,
                        int total_num_v)
{
    clear_sched();
    std::vector<float> *vv = val ? &g_val : NULL;
    switch (kind) {
        case 0:
            locality_schedule(ptr, idx, par_num, num_v, &g_ptr, &g_idx, &g_target, total_num_v, val, vv);
            break;
        case 1:
            neighbor_grouping_schedule(ptr, idx, neighbor_num, num_v, num_e, &g_ptr, &g_idx, &g_target);
            break;
        case 2:
            localityNeighborGrouping(ptr, idx, par_num, neighbor_num, num_v, &g_ptr, &g_idx, &g_target, total_num_v,
                                     val, vv);
            break;
        default:
            return -1;
    }
    return (long long)g_target.size();
}

long long ref_sched_num_edges() { return (long long)g_idx.size(); }
long long ref_sched_num_ptr() { return (long long)g_ptr.size(); }

void ref_sched_fetch(int *out_ptr, int *out_idx, int *out_target, float *out_val)
{
    if (out_ptr) memcpy(out_ptr, g_ptr.data(), g_ptr.size() * sizeof(int));
    if (out_idx) memcpy(out_idx, g_idx.data(), g_idx.size() * sizeof(int));
    if (out_target) memcpy(out_target, g_target.data(), g_target.size() * sizeof(int));
    if (out_val) memcpy(out_val, g_val.data(), g_val.size() * sizeof(float));
}

void ref_reorder_csr(const int *ptr, const int *idx, const int *map, const int *reverse_map, int num_v, int num_e,
                     int *newptr, int *newidx)
{
    int *p = newptr, *q = newidx;
    reorderCSR(ptr, idx, map, reverse_map, num_v, num_e, p, q);
}

// load_graph reads ../data/<dset>.* relative to the CWD (src/data.cu:34).  Returns arrays the
// caller copies out with ref_load_fetch().
static int *l_ptr = NULL, *l_idx = NULL;
static int l_nv = 0, l_ne = 0;
int ref_load_graph(const char *dset, const char *reorder_subfix, int *num_v, int *num_e)
{
    l_ptr = l_idx = NULL;
    rows = reverse_rows = NULL;
    reorderfile = "";
    load_graph(std::string(dset), l_nv, l_ne, l_ptr, l_idx, true, std::string(reorder_subfix));
    *num_v = l_nv;
    *num_e = l_ne;
    return 0;
}
void ref_load_fetch(int *out_ptr, int *out_idx, int *out_rows, int *out_reverse_rows)
{
    memcpy(out_ptr, l_ptr, (size_t)(l_nv + 1) * sizeof(int));
    memcpy(out_idx, l_idx, (size_t)l_ne * sizeof(int));
    if (rows && out_rows) memcpy(out_rows, rows, (size_t)l_nv * sizeof(int));
    if (reverse_rows && out_reverse_rows) memcpy(out_reverse_rows, reverse_rows, (size_t)l_nv * sizeof(int));
}
int ref_load_was_reordered() { return rows != NULL; }

// ---------------------------------------------------------------- device: the reference aggregators
void ref_set_globals(int nn, int mm)
{
    n = nn;
    m = mm;
}

// all device pointers are borrowed: register them so the reference never cudaFree()s them
void *ref_gcn_create(int *d_ptr, int *d_idx, float *d_val, int num_v, int num_e, int feat_in, int feat_out)
{
    registerPtr(d_ptr);
    registerPtr(d_idx);
    registerPtr(d_val);
    return new Aggregator_GCN(NULL, NULL, d_ptr, d_idx, num_v, num_e, feat_in, feat_out, d_val);
}
void ref_gcn_schedule(void *h, int kind, int p0, int p1)
{
    int arr[2] = {p0, p1};
    ((Aggregator_GCN *)h)->schedule((Schedule)kind, arr);
}
int ref_num_target(void *h) { return ((Aggregator *)h)->num_target; }
void ref_gcn_run(void *h, float *x, float *y, int block, int scheduled, int feat)
{
    ((Aggregator_GCN *)h)->feat_out = feat;  // the memset at aggr_gcn.h:427 uses feat_out
    ((Aggregator_GCN *)h)->run_with_feat(x, y, block, scheduled != 0, feat);
}
void ref_gcn_updateval(void *h, float *val)
{
    registerPtr(val);
    ((Aggregator_GCN *)h)->updateval(val);
}
void ref_gcn_run_with_nn(void *h, float *x, float *vout, float *w, float *transformed, int block)
{
    ((Aggregator_GCN *)h)->run_with_nn(x, vout, w, transformed, block);
}
double ref_gcn_run_edgewise(void *h, float *x, float *y, int block)
{
    return ((Aggregator_GCN *)h)->runEdgeWise(x, y, block, false);
}

void *ref_gat_create(int *d_ptr, int *d_idx, int num_v, int num_e, int feat)
{
    registerPtr(d_ptr);
    registerPtr(d_idx);
    return new Aggregator_GAT(NULL, NULL, d_ptr, d_idx, num_v, num_e, feat, feat);
}
void ref_gat_schedule(void *h, int kind, int p0, int p1)
{
    int arr[2] = {p0, p1};
    ((Aggregator_GAT *)h)->schedule((Schedule)kind, arr);
}
void ref_gat_run(void *h, float *x, float *att, float *y, int block, int scheduled, int feat)
{
    ((Aggregator_GAT *)h)->run_with_feat(x, att, y, block, scheduled != 0, feat);
}
void ref_gat_run_att(void *h, float *att, float *out_val, int block) { ((Aggregator_GAT *)h)->run_att(att, out_val, block); }
void ref_gat_run_u_add_v(void *h, float *att, float *out_val, int block)
{
    ((Aggregator_GAT *)h)->run_u_add_v(att, out_val, block);
}
void ref_gat_run_add_to_center(void *h, float *in_val, float *out_att, int block)
{
    ((Aggregator_GAT *)h)->run_add_to_center(in_val, out_att, block);
}
void ref_gat_run_div_each(void *h, float *in_att, float *inout_val, int block)
{
    ((Aggregator_GAT *)h)->run_div_each(in_att, inout_val, block);
}

// the experimental backward (aggr_gat.h:426-434); needs a neighbor_grouping schedule (it walks the scheduled arrays)
void ref_gat_run_bwd(void *h, float *output, float *doutput, float *newval, float *div, float *infeat, float *d_a_b,
                     float *d_feat, float relu_l, int block)
{
    ((Aggregator_GAT *)h)->run_bwd(output, doutput, newval, div, infeat, d_a_b, d_feat, relu_l, block);
}

// sampleVertex (sample.h:131-200) on device arrays; outputs are the reference's own cudaMalloc2'ed arrays
int ref_sample_vertex(int *d_active, int *d_ptr, int *d_idx, int layer_num, int **vertexset, int **sub_ptr, int **sub_idx,
                      int *num_e)
{
    CSRSubGraph g = sampleVertex(d_active, d_ptr, d_idx, layer_num);
    *vertexset = g.vertexset;
    *sub_ptr = g.ptr;
    *sub_idx = g.idx;
    *num_e = g.num_e;
    return g.num_v;
}

void *ref_sddmm_create(int *d_ptr, int *d_idx, int num_v, int num_e, int feat)
{
    registerPtr(d_ptr);
    registerPtr(d_idx);
    return new Aggregator_SDDMM(NULL, NULL, d_ptr, d_idx, num_v, num_e, feat, feat);
}
void ref_sddmm_schedule(void *h, int kind, int p0, int p1)
{
    int arr[2] = {p0, p1};
    ((Aggregator_SDDMM *)h)->schedule((Schedule)kind, arr);
}
double ref_sddmm_run(void *h, float *v1, float *v2, float *outval, int block, int scheduled)
{
    return ((Aggregator_SDDMM *)h)->run(v1, v2, outval, block, scheduled != 0);
}

void *ref_mlp_create(int *d_ptr, int *d_idx, int num_v, int num_e, float *d_weight)
{
    registerPtr(d_ptr);
    registerPtr(d_idx);
    return new Aggregator_MLP(NULL, NULL, d_ptr, d_idx, num_v, num_e, 32, 32, d_weight);  // aggr_nn.h is F = 32 only
}
void ref_mlp_schedule(void *h, int kind, int p0, int p1)
{
    int arr[2] = {p0, p1};
    ((Aggregator_MLP *)h)->schedule((Schedule)kind, arr);
}
double ref_mlp_run(void *h, float *vin, float *vout, int block, int scheduled)
{
    return ((Aggregator_MLP *)h)->run(vin, vout, block, scheduled != 0);
}

// un-fused combination baseline (dense.h:4-23); tmp is an [M,N] scratch
void ref_matmul_NN(float *A, float *B, float *C, int M, int N, int K, float *tmp)
{
    static bool created = false;  // cublasHs[] is an uninitialised new[] in src/util.cu:174
    if (!created) {
        cublasCreate(&cublasHs[0]);
        created = true;
    }
    matmul_NN(A, B, C, M, N, K, tmp);
}

int ref_sync() { return (int)cudaDeviceSynchronize(); }
}
