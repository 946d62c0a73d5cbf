// aggr_sddmm.h -- compatibility layer: class Aggregator_SDDMM of the reference
// (include/aggr_sddmm.h:85-120) on top of libgnnagg.so.
#ifndef AGGR_SDDMM_H
#define AGGR_SDDMM_H
#include "aggregator.h"

class Aggregator_SDDMM : public Aggregator {
public:
    Aggregator_SDDMM(int *host_out_ptr, int *host_out_idx, int *dev_out_ptr, int *dev_out_idx, int out_num_v,
                     int out_num_e, int out_feat_in, int out_feat_out)
        : Aggregator(host_out_ptr, host_out_idx, dev_out_ptr, dev_out_idx, out_num_v, out_num_e, out_feat_in,
                     out_feat_out)
    {
    }
    Aggregator_SDDMM(CSRSubGraph g, int out_feat_in, int out_feat_out) : Aggregator(g, out_feat_in, out_feat_out) {}

    // outval[e] = <v1[idx[e],:], v2[row,:]>; self-timed like the reference (:104-116)
    double run(float *v1, float *v2, float *outval, int BLOCK_SIZE, bool scheduled) override
    {
        if (scheduled) assert(sche == neighbor_grouping);
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t0);
        checkGnnagg(gnnagg_sddmm(handle, v1, v2, outval, feat_in, scheduled, NULL));
        checkCudaErrors(cudaDeviceSynchronize());
        timestamp(t1);
        return getDuration(t0, t1);
    }
};
#endif
