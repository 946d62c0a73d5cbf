// aggr_nn.h -- compatibility layer: class Aggregator_MLP of the reference (include/aggr_nn.h:290-341), the
// per-edge MLP aggregator  out[v] = sum_u ReLU((in[v] + in[u]) * W),  on top of libgnnagg.so
// (gnnagg_mlp_run: projection hoisted to one tensor-core GEMM, then a gather/ReLU/sum aggregation).
#ifndef AGGR_NN_H
#define AGGR_NN_H
#include "aggregator.h"

class Aggregator_MLP : public Aggregator {
public:
    Aggregator_MLP(int *host_out_ptr, int *host_out_idx, int *dev_out_ptr, int *dev_out_idx, int out_num_v,
                   int out_num_e, int out_feat_in, int out_feat_out, float *out_weight)
        : Aggregator(host_out_ptr, host_out_idx, dev_out_ptr, dev_out_idx, out_num_v, out_num_e, out_feat_in,
                     out_feat_out),
          d_weight(out_weight)
    {
    }
    Aggregator_MLP(CSRSubGraph g, int out_feat_in, int out_feat_out, float *out_weight)
        : Aggregator(g, out_feat_in, out_feat_out), d_weight(out_weight)
    {
    }
    // self-timed like the reference (:318-337): synchronises and returns seconds
    double run(float *vin, float *vout, int BLOCK_SIZE, bool scheduled) override
    {
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t0);
        checkGnnagg(gnnagg_mlp_run(handle, vin, d_weight, vout, feat_in, scheduled, NULL));
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t1);
        return getDuration(t0, t1);
    }

private:
    float *d_weight;
};
#endif
