// aggr_gcn.h -- compatibility layer: class Aggregator_GCN of the reference (include/aggr_gcn.h:362-550)
// on top of libgnnagg.so.  The kernels live in the library (gnn-computing_b200/csrc).
#ifndef AGGR_GCN_H
#define AGGR_GCN_H
#include "aggregator.h"

typedef uint64_t clocktype;

class Aggregator_GCN : public Aggregator {
public:
    Aggregator_GCN(int *host_out_ptr, int *host_out_idx, int *dev_out_ptr, int *dev_out_idx, int out_num_v,
                   int out_num_e, int out_feat_in, int out_feat_out, float *out_val)
        : Aggregator(host_out_ptr, host_out_idx, dev_out_ptr, dev_out_idx, out_num_v, out_num_e, out_feat_in,
                     out_feat_out),
          d_val(out_val)
    {
        checkGnnagg(gnnagg_set_val(handle, d_val));
    }
    Aggregator_GCN(CSRSubGraph g, int out_feat_in, int out_feat_out, float *out_val)
        : Aggregator(g, out_feat_in, out_feat_out), d_val(out_val)
    {
        checkGnnagg(gnnagg_set_val(handle, d_val));
    }
    ~Aggregator_GCN() { safeFree(d_val); }  // aggr_gcn.h:375-378

    // Y = A*X; asynchronous on the legacy default stream, returns 0.0 like the reference (:379-410)
    double run(float *vin, float *vout, int BLOCK_SIZE, bool scheduled) override
    {
        checkGnnagg(gnnagg_gcn_run(handle, vin, vout, feat_in, scheduled, NULL));
        return 0.0;
    }
    double run_with_feat(float *vin, float *vout, int BLOCK_SIZE, bool scheduled, int feat)
    {
        feat_in = feat;  // :413
        return run(vin, vout, BLOCK_SIZE, scheduled);
    }
    // self-timed like the reference (:445-460): synchronises and returns seconds
    double runEdgeWise(float *vin, float *vout, int BLOCK_SIZE, bool scheduled) override
    {
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t0);
        checkGnnagg(gnnagg_gcn_run_edgewise(handle, vin, vout, feat_in, NULL));
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t1);
        return getDuration(t0, t1);
    }
    // The reference's *_clock kernels (:159-248) store per-block (globaltimer begin, end, smid) for the
    // Figure 8 load-imbalance study; profiling is done with ncu here, so `timer` is left untouched and
    // only the self-timed duration is returned (:462-489).
    double run_clock(float *vin, float *vout, clocktype *timer, int BLOCK_SIZE, bool scheduled)
    {
        (void)timer;
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t0);
        run(vin, vout, BLOCK_SIZE, scheduled);
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t1);
        return getDuration(t0, t1);
    }
    // fused aggregation + combination (:491-499): vout = A*vin, transformed = vout*weight with
    // weight row-major [feat_in, feat_out]; requires a schedule like the reference.  Outputs are
    // overwritten (the reference accumulates into whatever the buffers held).
    void run_with_nn(float *vin, float *vout, float *weight, float *transformed, int BLOCK_SIZE)
    {
        checkGnnagg(gnnagg_gcn_layer(handle, vin, weight, transformed, vout, feat_in, feat_out, 1, NULL));
    }
    void schedule(Schedule s, int *param) override { Aggregator::schedule(s, param); }  // val handled inside the library
    void updateval(float *out_d_val)
    {
        d_val = out_d_val;
        checkGnnagg(gnnagg_set_val(handle, d_val));
    }

private:
    float *d_val = NULL;
};
#endif
