// aggr_gat.h -- compatibility layer: class Aggregator_GAT of the reference (include/aggr_gat.h:299-441)
// on top of libgnnagg.so.  LeakyReLU slope fixed at 0.2 as in the reference (:339,347,400).
#ifndef AGGR_GAT_H
#define AGGR_GAT_H
#include "aggregator.h"

class Aggregator_GAT : public Aggregator {
public:
    Aggregator_GAT(int *host_out_ptr, int *host_out_idx, int *dev_out_ptr, int *dev_out_idx, int out_num_v,
                   int out_num_e, int out_feat_in, int out_feat_out)
        : Aggregator(host_out_ptr, host_out_idx, dev_out_ptr, dev_out_idx, out_num_v, out_num_e, out_feat_in,
                     out_feat_out)
    {
    }
    Aggregator_GAT(CSRSubGraph g, int out_feat_in, int out_feat_out) : Aggregator(g, out_feat_in, out_feat_out) {}

    // fused SDDMM -> edge softmax -> SpMM (:317-354); vatt is the [n,2] attention table
    double run(float *vin, float *vatt, float *vout, int BLOCK_SIZE, bool scheduled) override
    {
        checkGnnagg(gnnagg_gat_run(handle, vin, vatt, vout, feat_in, 0.2f, scheduled, NULL));
        return 0.0;
    }
    double run_with_feat(float *vin, float *vatt, float *vout, int BLOCK_SIZE, bool scheduled, int feat)
    {
        feat_in = feat;
        return run(vin, vatt, vout, BLOCK_SIZE, scheduled);
    }
    void run_att(float *in_att, float *out_val, int BLOCK_SIZE)
    {
        checkGnnagg(gnnagg_edge_softmax(handle, in_att, out_val, 0.2f, NULL));
    }
    void run_u_add_v(float *in_att, float *out_val, int BLOCK_SIZE)
    {
        checkGnnagg(gnnagg_u_add_v(handle, in_att, out_val, NULL));
    }
    void run_add_to_center(float *in_val, float *out_att, int BLOCK_SIZE)
    {
        checkGnnagg(gnnagg_add_to_center(handle, in_val, out_att, NULL));
    }
    void run_div_each(float *in_att, float *in_out_val, int BLOCK_SIZE)
    {
        checkGnnagg(gnnagg_each_div(handle, in_att, in_out_val, NULL));
    }
    // backward of the scheduled fused aggregation (run_bwd + aggr_gat_fine_bwd, :222-294, :426-434): output = the
    // normalised forward result, newval = the un-normalised edge weights aggr_gat_fine left behind (:193), div = their
    // row sums (`scalar`).  Differences from the experimental reference kernel, all documented in gnnagg.h: any
    // feat_in (there 32 only), correct LeakyReLU derivative, BOTH halves of d_a_b, deterministic, and the two
    // outputs are overwritten rather than accumulated into caller-zeroed arrays.
    void run_bwd(float *output, float *doutput, float *newval, float *div, float *infeat, float *d_a_b, float *d_feat,
                 float relu_l, int BLOCK_SIZE)
    {
        if (!transposed) {
            checkGnnagg(gnnagg_transpose_build(handle, vertex_count(), NULL));
            transposed = true;
        }
        // newval arrives in SCHEDULED edge order (aggr_gat_fine wrote it that way); the backward indexes weights in CSR
        // order.  The two coincide for nop / neighbor_grouping; after a locality schedule the weights are mapped back.
        const int kind = gnnagg_schedule_kind(handle);
        float *w_csr = newval;
        if (kind == GNNAGG_SCHED_LOCALITY || kind == GNNAGG_SCHED_LOCALITY_NEIGHBOR_GROUPING) {
            if (!bwd_w) checkCudaErrors(cudaMalloc((void **)&bwd_w, sizeof(float) * (size_t)(edge_count() > 0 ? edge_count() : 1)));
            checkGnnagg(gnnagg_sched_to_csr_order(handle, newval, bwd_w, NULL));
            w_csr = bwd_w;
        }
        checkGnnagg(gnnagg_gat_backward(handle, infeat, NULL, w_csr, div, output, doutput, d_feat, d_a_b, feat_in, relu_l, NULL));
    }
    ~Aggregator_GAT()
    {
        if (bwd_w) cudaFree(bwd_w);
    }

private:
    bool transposed = false;
    float *bwd_w = NULL;  // newval in CSR order (run_bwd after a locality schedule)
};
#endif
