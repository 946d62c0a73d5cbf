// aggregator.h -- compatibility layer: class Aggregator of the reference (include/aggregator.h:25-151) as a
// thin C++ owner of a gnnagg_aggregator handle (C ABI: include/gnnagg.h).  The constructors, virtual methods and
// public fields a caller of the reference sees are kept; there are no kernels in this header -- every method
// forwards to libgnnagg.so.  BLOCK_SIZE arguments are accepted and ignored (launch shapes are internal).
#ifndef AGGREGATOR_H
#define AGGREGATOR_H

#include "data.h"
#include "graph_schedule.h"
#include "util.h"

class Aggregator {
    // --- what the reference exposes publicly (aggregator.h:124-126) ---
public:
    int feat_in = 0;
    int feat_out = 0;
    int num_target = 0;

    // device CSR (row = destination); host copies are optional and only forwarded to the library, which mirrors
    // the device arrays itself when a host-side pass needs them (reference: aggregator.h:30-39)
    Aggregator(int *host_ptr, int *host_idx, int *dev_ptr, int *dev_idx, int vertices, int edges, int fin, int fout)
        : feat_in(fin), feat_out(fout), csr_ptr_(dev_ptr), csr_idx_(dev_idx), mirror_ptr_(host_ptr),
          mirror_idx_(host_idx), vertices_(vertices), edges_(edges)
    {
        attach();
    }

    Aggregator(CSRSubGraph g, int fin, int fout)
        : feat_in(fin), feat_out(fout), csr_ptr_(g.ptr), csr_idx_(g.idx), vertex_set_(g.vertexset), vertices_(g.num_v),
          edges_(g.num_e)
    {
        attach();
    }

    // The reference aggregator takes over the device arrays it was constructed with and releases them unless they
    // were registerPtr()-ed (aggregator.h:58-66); callers rely on that, so the shim does the same.  The library
    // handle itself only ever frees what it allocated.
    virtual ~Aggregator()
    {
        gnnagg_destroy(handle);
        handle = NULL;
        safeFree(csr_ptr_);
        safeFree(csr_idx_);
        safeFree(vertex_set_);
        safeFree(edge_list_);
    }

    // param[0] = number of locality slices or neighbour-group size; param[1] = group size of the combined schedule.
    // Slices divide the global vertex count `n` (aggregator.h:79,87).  Built on the GPU by the library.
    virtual void schedule(Schedule s, int *param)
    {
        const int count = (s == locality_neighbor_grouping) ? 2 : 1;
        checkGnnagg(gnnagg_schedule_apply(handle, static_cast<int>(s), param, count, n));
        sche = s;
        switch (s) {
            case neighbor_grouping:
                neighbor_group_size = param[0];
                break;
            case locality_neighbor_grouping:
                neighbor_group_size = param[1];  // fall through: also a locality schedule
            case locality:
                locality_partition_num = param[0];
                break;
            default:
                break;
        }
        num_target = gnnagg_num_target(handle);
        dbg(num_target);
    }

    // the base class has no kernel of its own; subclasses override what they support (aggregator.h:100-114)
    virtual double run(float *, float *, int, bool) { return unsupported(); }
    virtual double run(float *, float *, float *, int, bool) { return unsupported(); }
    virtual double runEdgeWise(float *, float *, int, bool) { return unsupported(); }

    // (src, dst) pair list of the CSR, kept on the device (aggregator.h:115-122)
    void csr2edgelist()
    {
        safeFree(edge_list_);
        checkCudaErrors(cudaMalloc2((void **)&edge_list_, sizeof(int) * 2 * (size_t)edges_));
        checkGnnagg(gnnagg_csr2edgelist(handle, edge_list_, NULL));
    }

protected:
    gnnagg_aggregator *handle = NULL;
    Schedule sche = nop;
    int neighbor_group_size = 0;
    int locality_partition_num = 0;

    int *device_edge_list() const { return edge_list_; }
    int vertex_count() const { return vertices_; }
    int edge_count() const { return edges_; }

private:
    void attach() { checkGnnagg(gnnagg_create(csr_ptr_, csr_idx_, mirror_ptr_, mirror_idx_, vertices_, edges_, &handle)); }
    static double unsupported()
    {
        assert(false && "this aggregator does not implement the requested run variant");
        return -1;
    }

    int *csr_ptr_ = NULL;
    int *csr_idx_ = NULL;
    int *mirror_ptr_ = NULL;
    int *mirror_idx_ = NULL;
    int *vertex_set_ = NULL;
    int *edge_list_ = NULL;
    int vertices_ = 0;
    int edges_ = 0;
};
#endif
