// aggregator.h -- compatibility layer: class Aggregator of the reference (include/aggregator.h:25-151)
// as a thin C++ owner of a gnnagg_aggregator handle.  Same constructors, virtuals and public
// fields; every method forwards to the C ABI (include/gnnagg.h).  BLOCK_SIZE arguments are
// accepted and ignored (launch shapes are internal to the library).
#ifndef AGGREGATOR_H
#define AGGREGATOR_H

#include "data.h"
#include "graph_schedule.h"
#include "util.h"

class Aggregator {
public:
    Aggregator(int *host_out_ptr, int *host_out_idx, int *dev_out_ptr, int *dev_out_idx, int out_num_v, int out_num_e,
               int out_feat_in, int out_feat_out)
        : feat_in(out_feat_in), feat_out(out_feat_out), d_ptr(dev_out_ptr), d_idx(dev_out_idx), h_ptr(host_out_ptr),
          h_idx(host_out_idx), num_v(out_num_v), num_e(out_num_e)
    {
        open();
    }
    Aggregator(CSRSubGraph g, int out_feat_in, int out_feat_out)
        : feat_in(out_feat_in), feat_out(out_feat_out), d_ptr(g.ptr), d_idx(g.idx), d_vset(g.vertexset), num_v(g.num_v),
          num_e(g.num_e)
    {
        open();
    }
    // like the reference (aggregator.h:58-66) the aggregator takes over the device CSR it was given
    // unless the pointers were registerPtr()-ed
    virtual ~Aggregator()
    {
        gnnagg_destroy(handle);
        safeFree(d_ptr);
        safeFree(d_idx);
        safeFree(d_vset);
        safeFree(d_edgelist);
    }
    virtual void schedule(Schedule s, int *param)
    {
        sche = s;
        const int np = (s == locality_neighbor_grouping) ? 2 : 1;
        checkGnnagg(gnnagg_schedule_apply(handle, (int)s, param, np, n));  // slices divide the global n (aggregator.h:79)
        if (s == locality || s == locality_neighbor_grouping) locality_partition_num = param[0];
        if (s == neighbor_grouping) neighbor_group_size = param[0];
        if (s == locality_neighbor_grouping) neighbor_group_size = param[1];
        num_target = gnnagg_num_target(handle);
        dbg(num_target);
    }
    virtual double run(float *vin, float *vout, int BLOCK_SIZE, bool scheduled)
    {
        assert(false);
        return -1;
    }
    virtual double run(float *v1, float *v2, float *outval, int BLOCK_SIZE, bool scheduled)
    {
        assert(false);
        return -1;
    }
    virtual double runEdgeWise(float *vin, float *vout, int BLOCK_SIZE, bool scheduled)
    {
        assert(false);
        return -1;
    }
    void csr2edgelist()
    {
        safeFree(d_edgelist);
        checkCudaErrors(cudaMalloc2((void **)&d_edgelist, 2 * (size_t)num_e * sizeof(int)));
        checkGnnagg(gnnagg_csr2edgelist(handle, d_edgelist, NULL));
    }

    int feat_in = 0;
    int feat_out = 0;
    int num_target = 0;

protected:
    void open() { checkGnnagg(gnnagg_create(d_ptr, d_idx, h_ptr, h_idx, num_v, num_e, &handle)); }

    gnnagg_aggregator *handle = NULL;
    int *d_ptr = NULL;
    int *d_idx = NULL;
    int *h_ptr = NULL;
    int *h_idx = NULL;
    int *d_vset = NULL;
    int *d_edgelist = NULL;
    int num_v = 0;
    int num_e = 0;
    int neighbor_group_size = 0;
    int locality_partition_num = 0;
    Schedule sche = nop;
};
#endif
