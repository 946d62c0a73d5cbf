/*
 * oracle.c -- CPU restatement of the GNN-Computing neighbour-aggregation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Every function restates one piece of the reference (paths relative to /root/reference):
 *   float path   : include/aggr_gcn.h, include/aggr_gat.h, include/aggr_sddmm.h, include/dense.h
 *   integer path : include/graph_schedule.h, src/data.cu
 * Parity pinning : the integer functions are checked against the reference's own host code
 *   (compiled from where it lies into oracle/_ref/libref.so, see oracle/Makefile) and against
 *   the golden vectors under tests/golden/ that were produced by that library.  The reference
 *   has NO CPU implementation of the float kernels and NO tests; the float functions restate
 *   the kernel arithmetic and are pinned on the GPU box against the reference's own kernels
 *   recompiled for sm_100 (oracle/_ref/libref.so, tests/test_gpu_reference_kernels.py).
 *
 * Layout conventions: CSR row = destination vertex, idx = source
 * neighbours in file order, int32 indices, fp32 values, dense matrices row-major [rows, F].
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is meant to use all host cores */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * GCN weighted-sum aggregation  Y[r,c] = sum_{e in row r} val[e] * X[idx[e], c]
 * reference: aggr_gcn.h:13-35 (kernel aggr_gcn): per (row, col) a sequential chain
 * `rs += vin[idx*F + col] * val` in CSR order; built with --use_fast_math so the chain is an
 * FFMA chain (CMakeLists.txt:40).  Empty rows produce 0 (aggr_gcn.h:14,35).
 * fp32 variant keeps that order with fmaf; it is the CPU-timing / "port" baseline.
 * ------------------------------------------------------------------------------------------ */
void orc_spmm_f32(int64_t row_begin, int64_t row_end, const int *ptr, const int *idx, const float *val,
                  const float *X, int F, float *Y)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = row_begin; r < row_end; ++r) {
        float *y = Y + (size_t)r * F;
        for (int c = 0; c < F; ++c) y[c] = 0.0f;
        for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
            const float *x = X + (size_t)idx[e] * F;
            const float w = val[e];
            for (int c = 0; c < F; ++c) y[c] = fmaf(x[c], w, y[c]);
        }
    }
}

/* fp64-accumulating ground truth of the same sum; also returns scale[r,c] = sum |val*x| that
 * the parity gate uses:  |y - y64| <= tol * scale  (SURVEY.md 8(d) "parity gate"). */
void orc_spmm_f64(int64_t n, const int *ptr, const int *idx, const float *val, const float *X, int F,
                  float *Y, float *scale)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n; ++r) {
        double *acc = (double *)calloc((size_t)2 * F, sizeof(double));
        double *mag = acc + F;
        for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
            const float *x = X + (size_t)idx[e] * F;
            const double w = val[e];
            for (int c = 0; c < F; ++c) {
                const double t = w * (double)x[c];
                acc[c] += t;
                mag[c] += fabs(t);
            }
        }
        for (int c = 0; c < F; ++c) {
            Y[(size_t)r * F + c] = (float)acc[c];
            if (scale) scale[(size_t)r * F + c] = (float)mag[c];
        }
        free(acc);
    }
}

/* Scheduled (grouped) aggregation: group g sums edges [gptr[g], gptr[g+1]) of (gidx, gval) and
 * adds the partial into row target[g] of a pre-zeroed Y.
 * reference: aggr_gcn.h:78-114 (kernel aggr_gcn_target, atomicAdd at :112) with the memset of
 * aggr_gcn.h:393.  Order of the cross-group combination is unspecified there; fp64 here. */
void orc_spmm_grouped_f64(int64_t n, int64_t num_groups, const int *gptr, const int *gidx, const float *gval,
                          const int *target, const float *X, int F, float *Y, float *scale)
{
    double *acc = (double *)calloc((size_t)n * F, sizeof(double));
    double *mag = (double *)calloc((size_t)n * F, sizeof(double));
    for (int64_t g = 0; g < num_groups; ++g) {
        double *a = acc + (size_t)target[g] * F;
        double *s = mag + (size_t)target[g] * F;
        for (int e = gptr[g]; e < gptr[g + 1]; ++e) {
            const float *x = X + (size_t)gidx[e] * F;
            const double w = gval[e];
            for (int c = 0; c < F; ++c) {
                const double t = w * (double)x[c];
                a[c] += t;
                s[c] += fabs(t);
            }
        }
    }
    for (size_t i = 0; i < (size_t)n * F; ++i) {
        Y[i] = (float)acc[i];
        if (scale) scale[i] = (float)mag[i];
    }
    free(acc);
    free(mag);
}

/* Edge-wise aggregation: for each (src,dst) pair  Y[dst] += X[src]*val[e]
 * reference: aggr_gcn.h:291-302 (aggr_gcn_edgewise) over the edge list produced by
 * aggregator.h:11-23 (convertCSRToEdgelist: edgelist[2e]=idx[e] (src), edgelist[2e+1]=row). */
void orc_csr2edgelist(int64_t n, const int *ptr, const int *idx, int *edgelist)
{
    for (int64_t r = 0; r < n; ++r)
        for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
            edgelist[2 * (size_t)e] = idx[e];
            edgelist[2 * (size_t)e + 1] = (int)r;
        }
}

/* ------------------------------------------------------------------------------------------
 * Dense combination  H = AX * W,  W row-major [IN, OUT]
 * reference: aggr_gcn.h:341-357 (aggr_gcn_nn: ans[lane] = sum_i W[i*OUT+lane]*rs_i) and the
 * un-fused baseline dense.h:4-23 (matmul_NN: row-major C[M,N] = A[M,K] * B[K,N]).
 * ------------------------------------------------------------------------------------------ */
void orc_dense_f64(int64_t M, int K, int N, const float *A, const float *W, float *H, float *scale)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        for (int o = 0; o < N; ++o) {
            double acc = 0.0, mag = 0.0;
            for (int k = 0; k < K; ++k) {
                const double t = (double)A[(size_t)i * K + k] * (double)W[(size_t)k * N + o];
                acc += t;
                mag += fabs(t);
            }
            H[(size_t)i * N + o] = (float)acc;
            if (scale) scale[(size_t)i * N + o] = (float)mag;
        }
    }
}

/* fp32 row-block variant of the combination, used only to TIME the CPU port (bench.py
 * cpu_baseline / --impl reference): same i-k-o loop order a scalar port would use. */
void orc_dense_f32(int64_t row_begin, int64_t row_end, int K, int N, const float *A, const float *W, float *H)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = row_begin; i < row_end; ++i) {
        float *h = H + (size_t)i * N;
        for (int o = 0; o < N; ++o) h[o] = 0.0f;
        for (int k = 0; k < K; ++k) {
            const float a = A[(size_t)i * K + k];
            const float *w = W + (size_t)k * N;
            for (int o = 0; o < N; ++o) h[o] = fmaf(a, w[o], h[o]);
        }
    }
}

/* Fused layer H = (A*X)*W in fp64 end to end (no fp32 rounding of the intermediate); scale is
 * sum over edges and k of |val*x*w| so the 1e-5 gate covers both stages.
 * reference semantics: aggr_gcn.h:304-359 + 491-499 (run_with_nn), Figure10/main_b.cu:84-101. */
void orc_gcn_layer_f64(int64_t n, const int *ptr, const int *idx, const float *val, const float *X, int K,
                       const float *W, int N, float *AX, float *H, float *scale)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n; ++r) {
        double *acc = (double *)calloc((size_t)2 * K, sizeof(double));
        double *mag = acc + K;
        for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
            const float *x = X + (size_t)idx[e] * K;
            const double w = val[e];
            for (int c = 0; c < K; ++c) {
                const double t = w * (double)x[c];
                acc[c] += t;
                mag[c] += fabs(t);
            }
        }
        if (AX)
            for (int c = 0; c < K; ++c) AX[(size_t)r * K + c] = (float)acc[c];
        for (int o = 0; o < N; ++o) {
            double h = 0.0, s = 0.0;
            for (int k = 0; k < K; ++k) {
                const double w = (double)W[(size_t)k * N + o];
                h += acc[k] * w;
                s += mag[k] * fabs(w);
            }
            H[(size_t)r * N + o] = (float)h;
            if (scale) scale[(size_t)r * N + o] = (float)s;
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------
 * GAT.  Attention table att fp32 [n,2]: att[2v] = destination term, att[2u+1] = source term.
 *   s = att[2v] + att[2u+1];  w = exp(max(s, slope*s))      (aggr_gat.h:12-17, 125-143)
 * No max-subtraction (aggr_gat.h:14-20).  The reference uses __expf / fast division; the oracle
 * uses exp() in fp64, tests allow for that.
 * ------------------------------------------------------------------------------------------ */
static inline double gat_w(const float *att, int64_t v, int u, float slope)
{
    const float s = att[2 * v] + att[2 * (size_t)u + 1]; /* fp32 add as on the GPU */
    const float l = s * slope;
    return exp((double)(s > l ? s : l));
}

/* fused GAT aggregation  Y[v] = sum_u w_uv X[u] / sum_u w_uv      (aggr_gat.h:116-164)
 * Empty rows: the reference stores 0/0 = NaN (aggr_gat.h:162-163); `empty_value` selects what the
 * oracle writes there so both the reference behaviour (NaN) and the product's documented 0 can
 * be checked.  den (nullable) receives sum_u w_uv, scale the |.|-sum of the normalised terms. */
void orc_gat_f64(int64_t n, const int *ptr, const int *idx, const float *att, float slope, const float *X, int F,
                 float empty_value, float *Y, float *den, float *scale)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t v = 0; v < n; ++v) {
        double *acc = (double *)calloc((size_t)2 * F, sizeof(double));
        double *mag = acc + F;
        double d = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const double w = gat_w(att, v, idx[e], slope);
            const float *x = X + (size_t)idx[e] * F;
            d += w;
            for (int c = 0; c < F; ++c) {
                acc[c] += w * (double)x[c];
                mag[c] += fabs(w * (double)x[c]);
            }
        }
        if (den) den[v] = (float)d;
        for (int c = 0; c < F; ++c) {
            if (ptr[v] == ptr[v + 1]) {
                Y[(size_t)v * F + c] = empty_value;
                if (scale) scale[(size_t)v * F + c] = 0.0f;
            } else {
                Y[(size_t)v * F + c] = (float)(acc[c] / d);
                if (scale) scale[(size_t)v * F + c] = (float)(mag[c] / d);
            }
        }
        free(acc);
    }
}

/* edge softmax  newval[e] = w_e / sum_{e' in row} w_e'          (aggr_gat.h:5-31, attGat) */
void orc_edge_softmax_f64(int64_t n, const int *ptr, const int *idx, const float *att, float slope, float *newval)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t v = 0; v < n; ++v) {
        double d = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) d += gat_w(att, v, idx[e], slope);
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) newval[e] = (float)(gat_w(att, v, idx[e], slope) / d);
    }
}

/* un-normalised edge weights, what aggr_gat_fine leaves in newval (aggr_gat.h:186-193) */
void orc_edge_weight_f64(int64_t n, const int *ptr, const int *idx, const float *att, float slope, float *newval)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t v = 0; v < n; ++v)
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) newval[e] = (float)gat_w(att, v, idx[e], slope);
}

/* un-fused pieces (Figure10/main_a.cu:84-86):
 *   u_add_v        newval[e] = att[2v] + att[2u+1]              (aggr_gat.h:33-48)
 *   add_to_center  out[v]    = sum_e newval[e]   (stride 1!)    (aggr_gat.h:50-74, :71)
 *   each_div       newval[e] /= in[v]            (stride 1)     (aggr_gat.h:76-92)      */
void orc_u_add_v(int64_t n, const int *ptr, const int *idx, const float *att, float *newval)
{
    for (int64_t v = 0; v < n; ++v)
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) newval[e] = att[2 * v] + att[2 * (size_t)idx[e] + 1];
}

void orc_add_to_center_f64(int64_t n, const int *ptr, const float *newval, float *out)
{
    for (int64_t v = 0; v < n; ++v) {
        double s = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) s += (double)newval[e];
        out[v] = (float)s;
    }
}

void orc_each_div(int64_t n, const int *ptr, const float *in, float *newval)
{
    for (int64_t v = 0; v < n; ++v)
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) newval[e] = newval[e] / in[v];
}

/* ------------------------------------------------------------------------------------------
 * per-edge MLP aggregation  Y[v,o] = sum_{u in N(v)} ReLU( sum_k (X[v,k] + X[u,k]) * W[k*F + o] )
 * reference: aggr_nn.h:11-48 (macro COMP, the PURE_COMP path both kernels use): input[k] = cached[k] +
 * vin[idx*32 + k] (:35), ans = sum_k input[k] * shared_weight[lane + 32*k] (:38-41), `if (ans > 0) rs += ans`
 * (:42-43); F = 32 there.  scale = sum over edges and k of |(x_v + x_u) w| for the parity gate.
 * ------------------------------------------------------------------------------------------ */
void orc_mlp_f64(int64_t n, const int *ptr, const int *idx, const float *X, const float *W, int F, float *Y,
                 float *scale)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t v = 0; v < n; ++v) {
        double *acc = (double *)calloc((size_t)2 * F, sizeof(double));
        double *mag = acc + F;
        const float *xv = X + (size_t)v * F;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const float *xu = X + (size_t)idx[e] * F;
            for (int o = 0; o < F; ++o) {
                double ans = 0.0, m = 0.0;
                for (int k = 0; k < F; ++k) {
                    const double t = ((double)xv[k] + (double)xu[k]) * (double)W[(size_t)k * F + o];
                    ans += t;
                    m += fabs((double)xv[k] * (double)W[(size_t)k * F + o]) + fabs((double)xu[k] * (double)W[(size_t)k * F + o]);
                }
                if (ans > 0) acc[o] += ans;
                mag[o] += m;
            }
        }
        for (int o = 0; o < F; ++o) {
            Y[(size_t)v * F + o] = (float)acc[o];
            if (scale) scale[(size_t)v * F + o] = (float)mag[o];
        }
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------
 * SDDMM  val[e] = < X1[idx[e], 0:F], X2[row, 0:F] >             (aggr_sddmm.h:17-41)
 * The reference hard-wires the X1 row stride to 32 (aggr_sddmm.h:21,27) and reads 32 lanes;
 * with F = 32 both agree, which is the only case the reference supports.
 * ------------------------------------------------------------------------------------------ */
void orc_sddmm_f64(int64_t n, const int *ptr, const int *idx, const float *X1, const float *X2, int F, float *val,
                   float *scale)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n; ++r) {
        const float *b = X2 + (size_t)r * F;
        for (int e = ptr[r]; e < ptr[r + 1]; ++e) {
            const float *a = X1 + (size_t)idx[e] * F;
            double acc = 0.0, mag = 0.0;
            for (int k = 0; k < F; ++k) {
                acc += (double)a[k] * (double)b[k];
                mag += fabs((double)a[k] * (double)b[k]);
            }
            val[e] = (float)acc;
            if (scale) scale[e] = (float)mag;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Schedules (integer, bit-exact contract).  Two-call protocol: pass NULL outputs to obtain the
 * number of groups, then call again with buffers of that size.
 * ------------------------------------------------------------------------------------------ */

/* neighbor grouping: graph_schedule.h:91-154.  Every row is cut into consecutive chunks of at
 * most `neighbor_num` edges (:103-111), a non-empty remainder closes the row (:112-119), empty
 * rows emit nothing, idx_vec is a verbatim copy of idx (:123-124). */
int64_t orc_neighbor_grouping(const int *ptr, const int *idx, int neighbor_num, int num_v, int num_e, int *out_ptr,
                              int *out_idx, int *out_target)
{
    int64_t g = 0;
    if (out_ptr) out_ptr[0] = 0;
    for (int i = 0; i < num_v; ++i) {
        int left = ptr[i];
        while (ptr[i + 1] - left > neighbor_num) {
            left += neighbor_num;
            if (out_ptr) {
                out_ptr[g + 1] = left;
                out_target[g] = i;
            }
            ++g;
        }
        if (ptr[i + 1] != left) {
            if (out_ptr) {
                out_ptr[g + 1] = ptr[i + 1];
                out_target[g] = i;
            }
            ++g;
        }
    }
    if (out_idx) memcpy(out_idx, idx, (size_t)num_e * sizeof(int));
    return g;
}

/* locality schedule (neighbor_num <= 0): graph_schedule.h:17-89
 * locality + neighbor grouping (neighbor_num > 0): graph_schedule.h:156-243
 * Source-id range [0,total_num_v) is cut into par_num slices of floor(total/par) ids, the last
 * slice runs to total_num_v (:26-29 / :165-168); for every slice, every row in order, the
 * neighbours inside the slice are appended in CSR order (:35-43); a group is closed per
 * (slice,row) with >=1 hit (:54-57) and, in the LNG variant, additionally every neighbor_num
 * hits (:182-190) with the remainder flushed per (slice,row) (:202-209).  val is permuted
 * alongside when given (:41-42). */
int64_t orc_locality(const int *ptr, const int *idx, const float *val, int par_num, int neighbor_num, int num_v,
                     int total_num_v, int *out_ptr, int *out_idx, int *out_target, float *out_val)
{
    int64_t g = 0, pos = 0;
    if (out_ptr) out_ptr[0] = 0;
    for (int par = 0; par < par_num; ++par) {
        const int llim = par * (total_num_v / par_num);
        int ulim = llim + total_num_v / par_num;
        if (par == par_num - 1) ulim = total_num_v;
        for (int i = 0; i < num_v; ++i) {
            int cnt = 0;
            for (int j = ptr[i]; j < ptr[i + 1]; ++j) {
                if (idx[j] >= llim && idx[j] < ulim) {
                    ++cnt;
                    if (out_idx) out_idx[pos] = idx[j];
                    if (out_val && val) out_val[pos] = val[j];
                    ++pos;
                    if (neighbor_num > 0 && cnt == neighbor_num) {
                        if (out_ptr) {
                            out_ptr[g + 1] = (int)pos;
                            out_target[g] = i;
                        }
                        ++g;
                        cnt = 0;
                    }
                }
            }
            if (cnt != 0) {
                if (out_ptr) {
                    out_ptr[g + 1] = (int)pos;
                    out_target[g] = i;
                }
                ++g;
            }
        }
    }
    return g;
}

/* reorder application: src/data.cu:4-29 (reorderCSR).  map[i] = old id of the vertex placed at
 * new position i, reverse_map[old] = new.  Rows are permuted, neighbour ids relabelled, the
 * within-row order of the OLD row is kept (:19-24). */
void orc_reorder_csr(const int *ptr, const int *idx, const int *map, const int *reverse_map, int num_v,
                     int *newptr, int *newidx)
{
    int begin = 0;
    newptr[0] = 0;
    for (int i = 0; i < num_v; ++i) {
        const int base = ptr[map[i]];
        const int range = ptr[map[i] + 1] - base;
        for (int j = 0; j < range; ++j) newidx[begin + j] = reverse_map[idx[base + j]];
        begin += range;
        newptr[i + 1] = begin;
    }
}

/* loader: src/data.cu:31-139 (load_graph).  <dir><dset>.config = "num_v num_e" (:38-44);
 * row pointers then indices come from the raw int32 caches <dset>.graph.ptrdump/.edgedump when
 * they exist (:50-54, :77-81) else from the whitespace-separated text <dset>.graph, in which
 * case the caches are written (:56-68, :83-93); ptr[num_v] must equal num_e (:69-74).
 * Returns 0 on success.  Reorder application is a separate call (orc_read_reorder +
 * orc_reorder_csr), as at :96-133 (entry k of the file = old id placed at new position k). */
int orc_load_config(const char *dir, const char *dset, int *num_v, int *num_e)
{
    char path[4096];
    snprintf(path, sizeof path, "%s%s.config", dir, dset);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int ok = fscanf(f, "%d", num_v) == 1 && fscanf(f, "%d", num_e) == 1;
    fclose(f);
    return ok ? 0 : -2;
}

int orc_load_graph(const char *dir, const char *dset, int num_v, int num_e, int *indptr, int *indices)
{
    char graph[4096], ptrdump[4200], edgedump[4200];
    snprintf(graph, sizeof graph, "%s%s.graph", dir, dset);
    snprintf(ptrdump, sizeof ptrdump, "%s.ptrdump", graph);
    snprintf(edgedump, sizeof edgedump, "%s.edgedump", graph);
    FILE *txt = NULL, *f;
    if ((f = fopen(ptrdump, "rb"))) {
        size_t got = fread(indptr, sizeof(int), (size_t)num_v + 1, f);
        fclose(f);
        if (got != (size_t)num_v + 1) return -3;
    } else {
        if (!(txt = fopen(graph, "r"))) return -4;
        for (int i = 0; i <= num_v; ++i)
            if (fscanf(txt, "%d", indptr + i) != 1) return -5;
        if (!(f = fopen(ptrdump, "wb"))) return -6;
        fwrite(indptr, sizeof(int), (size_t)num_v + 1, f);
        fclose(f);
    }
    if (indptr[num_v] != num_e) return -7;
    if ((f = fopen(edgedump, "rb"))) {
        size_t got = fread(indices, sizeof(int), (size_t)num_e, f);
        fclose(f);
        if (txt) fclose(txt);
        if (got != (size_t)num_e) return -8;
    } else {
        /* the reference continues reading the SAME stream, i.e. it only works when the ptr line
         * was parsed from text in this call (src/data.cu:85 uses `fin` left open at :59) */
        if (!txt) return -9;
        for (int i = 0; i < num_e; ++i)
            if (fscanf(txt, "%d", indices + i) != 1) return -10;
        fclose(txt);
        if (!(f = fopen(edgedump, "wb"))) return -11;
        fwrite(indices, sizeof(int), (size_t)num_e, f);
        fclose(f);
    }
    return 0;
}

int orc_read_reorder(const char *path, int num_v, int *rows, int *reverse_rows)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    for (int i = 0; i < num_v; ++i) {
        int row;
        if (fscanf(f, "%d", &row) != 1) {
            fclose(f);
            return -2;
        }
        rows[i] = row;          /* src/data.cu:109-112 */
        reverse_rows[row] = i;
    }
    fclose(f);
    return 0;
}

/* mismatch counters: spmm.h:11-21 (validate2: relative error > 1e-2) and spmm.h:23-33
 * (validateReordered: absolute error > 1e-2 through the row map). */
int orc_validate2(const float *ref, const float *ans, int64_t num)
{
    int diff = 0;
    for (int64_t i = 0; i < num; ++i)
        if (fabsf((ref[i] - ans[i]) / ref[i]) > 1e-2f) ++diff;
    return diff;
}

int orc_validate_reordered(const float *ref, const float *ans, const int *map, int num_v, int F)
{
    int diff = 0;
    for (int64_t t = 0; t < (int64_t)num_v * F; ++t)
        if (fabsf(ref[t] - ans[(size_t)map[t / F] * F + t % F]) > 1e-2f) ++diff;
    return diff;
}

/* ------------------------------------------------------------------------------------------
 * Backward of the aggregation.  The reference only has the experimental
 * aggr_gat_fine_bwd (aggr_gat.h:222-294, run_bwd :426-434), used by no driver: F = 32 only
 * (`doutput[which_v*INFEATURE + lane]`, :245), the LeakyReLU derivative is keyed on
 * `newval < 0` which never holds for newval = exp(.) (:289), and only the source half of the
 * attention gradient is produced (:290).  What follows is the full derivative of the forward
 * oracle above; it coincides with that kernel where the kernel is right (F = 32, all pre-
 * activations positive, source half) -- tests/test_backward_gpu.py checks exactly that on the
 * GPU, tests/test_backward.py checks it against finite differences of orc_gat_f64.
 * ------------------------------------------------------------------------------------------ */

/* stable transpose of a CSR: edges grouped by source, inside a source in CSR order (hence by
 * ascending destination).  t_ptr[num_src+1], t_idx[m] = destination rows, t_perm[m] = CSR edge id. */
void orc_transpose_csr(int64_t n, const int *ptr, const int *idx, int num_src, int *t_ptr, int *t_idx, int *t_perm)
{
    const int m = ptr[n];
    memset(t_ptr, 0, ((size_t)num_src + 1) * sizeof(int));
    for (int e = 0; e < m; ++e) ++t_ptr[idx[e] + 1];
    for (int u = 0; u < num_src; ++u) t_ptr[u + 1] += t_ptr[u];
    int *fill = (int *)malloc(((size_t)num_src + 1) * sizeof(int));
    memcpy(fill, t_ptr, ((size_t)num_src + 1) * sizeof(int));
    for (int64_t v = 0; v < n; ++v)
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const int pos = fill[idx[e]]++;
            t_idx[pos] = (int)v;
            t_perm[pos] = e;
        }
    free(fill);
}

/* dX[u,:] = sum over edges (v <- u) of val_e * dY[v,:]   (gradient of orc_spmm w.r.t. X); fp64 */
void orc_spmm_t_f64(int64_t n, const int *ptr, const int *idx, const float *val, const float *dY, int F, int num_src,
                    float *dX, float *scale)
{
    double *acc = (double *)calloc((size_t)num_src * F * 2, sizeof(double));
    double *mag = acc + (size_t)num_src * F;
    for (int64_t v = 0; v < n; ++v)
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const double w = val[e];
            double *a = acc + (size_t)idx[e] * F, *g = mag + (size_t)idx[e] * F;
            const float *d = dY + (size_t)v * F;
            for (int c = 0; c < F; ++c) {
                a[c] += w * (double)d[c];
                g[c] += fabs(w * (double)d[c]);
            }
        }
    for (size_t i = 0; i < (size_t)num_src * F; ++i) {
        dX[i] = (float)acc[i];
        if (scale) scale[i] = (float)mag[i];
    }
    free(acc);
}

/* Backward of orc_gat_f64 for a loss with dL/dY = dY:
 *   alpha_e = w_e / D_v,  g_e = <X[u], dY[v]>,  c_v = <Y[v], dY[v]>
 *   ds_e    = alpha_e (g_e - c_v) * (s_e > 0 ? 1 : slope)            [= aggr_gat.h:287-289 where that is right]
 *   dX[u]  += alpha_e dY[v]                                          [:264]
 *   datt[2v] += ds_e (destination term), datt[2u+1] += ds_e (source term, :290)
 * att and datt have `att_rows` = max(n, num_src) rows.  dX_scale / datt_scale receive the sums of the
 * magnitudes of the terms (error bounds for the fp32 implementation). */
void orc_gat_backward_f64(int64_t n, const int *ptr, const int *idx, const float *att, float slope, const float *X,
                          const float *dY, int F, int num_src, int64_t att_rows, float *dX, float *datt, float *dX_scale,
                          float *datt_scale)
{
    double *ax = (double *)calloc((size_t)num_src * F * 2, sizeof(double));
    double *mx = ax + (size_t)num_src * F;
    double *da = (double *)calloc((size_t)att_rows * 4, sizeof(double));
    double *ma = da + (size_t)att_rows * 2;
    double *y = (double *)malloc((size_t)F * sizeof(double));
    for (int64_t v = 0; v < n; ++v) {
        if (ptr[v] == ptr[v + 1]) continue;
        const float *d = dY + (size_t)v * F;
        double D = 0.0;
        for (int c = 0; c < F; ++c) y[c] = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const double w = gat_w(att, v, idx[e], slope);
            const float *x = X + (size_t)idx[e] * F;
            D += w;
            for (int c = 0; c < F; ++c) y[c] += w * (double)x[c];
        }
        double cv = 0.0, cmag = 0.0;
        for (int c = 0; c < F; ++c) {
            cv += y[c] / D * (double)d[c];
            cmag += fabs(y[c] / D * (double)d[c]);
        }
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const int u = idx[e];
            const double alpha = gat_w(att, v, u, slope) / D;
            const float *x = X + (size_t)u * F;
            double g = 0.0, gmag = 0.0;
            for (int c = 0; c < F; ++c) {
                g += (double)x[c] * (double)d[c];
                gmag += fabs((double)x[c] * (double)d[c]);
                ax[(size_t)u * F + c] += alpha * (double)d[c];
                mx[(size_t)u * F + c] += fabs(alpha * (double)d[c]);
            }
            const float s = att[2 * v] + att[2 * (size_t)u + 1];
            const double lr = (s > 0.0f) ? 1.0 : (double)slope;
            const double ds = alpha * (g - cv) * lr, dm = alpha * (gmag + cmag) * lr;
            da[2 * v] += ds, ma[2 * v] += dm;
            da[2 * (size_t)u + 1] += ds, ma[2 * (size_t)u + 1] += dm;
        }
    }
    for (size_t i = 0; i < (size_t)num_src * F; ++i) {
        dX[i] = (float)ax[i];
        if (dX_scale) dX_scale[i] = (float)mx[i];
    }
    for (size_t i = 0; i < (size_t)att_rows * 2; ++i) {
        datt[i] = (float)da[i];
        if (datt_scale) datt_scale[i] = (float)ma[i];
    }
    free(ax), free(da), free(y);
}

/* scalar loss L = sum(dY * Y) of the forward oracle, all in fp64 with fp64 inputs: the function the
 * finite-difference test differentiates (tests/test_backward.py) */
double orc_gat_loss_f64(int64_t n, const int *ptr, const int *idx, const double *att, double slope, const double *X,
                        const double *dY, int F)
{
    double L = 0.0;
    double *y = (double *)malloc((size_t)F * sizeof(double));
    for (int64_t v = 0; v < n; ++v) {
        if (ptr[v] == ptr[v + 1]) continue;
        double D = 0.0;
        for (int c = 0; c < F; ++c) y[c] = 0.0;
        for (int e = ptr[v]; e < ptr[v + 1]; ++e) {
            const double s = att[2 * v] + att[2 * (size_t)idx[e] + 1];
            const double w = exp(s > s * slope ? s : s * slope);
            D += w;
            for (int c = 0; c < F; ++c) y[c] += w * X[(size_t)idx[e] * F + c];
        }
        for (int c = 0; c < F; ++c) L += y[c] / D * dY[(size_t)v * F + c];
    }
    free(y);
    return L;
}

/* ------------------------------------------------------------------------------------------
 * Sub-graph samplers (include/sample.h).  orc_sample_subgraph with fanout <= 0 restates sampleVertex
 * (:131-200): expandActive hops (:109-124) then rows of the active vertices in ascending order with
 * complete neighbour lists.  With fanout > 0 it is the SPECIFICATION of the product's fixed-fanout
 * sampler (gnn-computing_b200/csrc/sample_device.cu explains why sampleVertexSampleNeighbor,
 * :274-357, is not reproduced): parity unpinned against the reference for that mode.
 * ------------------------------------------------------------------------------------------ */
static uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int orc_sample_pos(uint64_t seed, int v, int j, int deg, int fanout)
{
    const int lo = (int)(((int64_t)j * deg) / fanout);
    const int hi = (int)(((int64_t)(j + 1) * deg) / fanout);
    const uint64_t r = orc_splitmix64(seed ^ ((uint64_t)(uint32_t)v * 0x9E3779B97F4A7C15ull) ^
                                      ((uint64_t)(uint32_t)j * 0xD1B54A32D192ED03ull));
    return lo + (int)(r % (uint64_t)(hi - lo));
}

/* active [n] in/out (0/1 after the call); vertexset [n], sub_ptr [n+1], sub_idx [cap] caller-allocated
 * (cap >= the number of extracted edges; m always suffices).  Returns the number of extracted rows,
 * *num_e the number of extracted edges, or -1 when sub_idx is too small. */
int orc_sample_subgraph(int n, const int *ptr, const int *idx, int *active, int fanout, int layer_num, uint64_t seed,
                        int *vertexset, int *sub_ptr, int *sub_idx, int64_t cap, int *num_e)
{
    int *next = (int *)malloc(((size_t)n + 1) * sizeof(int));
    for (int v = 0; v < n; ++v) active[v] = active[v] != 0;
    for (int hop = 0; hop < layer_num - 1; ++hop) {
        memcpy(next, active, (size_t)n * sizeof(int));
        for (int v = 0; v < n; ++v) {
            if (!active[v]) continue;
            const int deg = ptr[v + 1] - ptr[v];
            if (fanout <= 0 || deg <= fanout)
                for (int e = ptr[v]; e < ptr[v + 1]; ++e) next[idx[e]] = 1;
            else
                for (int j = 0; j < fanout; ++j) next[idx[ptr[v] + orc_sample_pos(seed, v, j, deg, fanout)]] = 1;
        }
        memcpy(active, next, (size_t)n * sizeof(int));
    }
    free(next);
    int rows = 0;
    int64_t edges = 0;
    sub_ptr[0] = 0;
    for (int v = 0; v < n; ++v) {
        if (!active[v]) continue;
        const int deg = ptr[v + 1] - ptr[v];
        const int take = (fanout > 0 && deg > fanout) ? fanout : deg;
        if (edges + take > cap) return -1;
        for (int j = 0; j < take; ++j)
            sub_idx[edges + j] = idx[ptr[v] + ((take == deg) ? j : orc_sample_pos(seed, v, j, deg, fanout))];
        edges += take;
        vertexset[rows] = v;
        sub_ptr[++rows] = (int)edges;
    }
    *num_e = (int)edges;
    return rows;
}
