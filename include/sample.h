// sample.h -- compatibility layer for the reference's include/sample.h: fullGraph (:126-129) and the two GPU
// samplers (sampleVertex :131-200, sampleVertexSampleNeighbor :274-357) on top of gnnagg_sample_subgraph.
// All pointers are device pointers; `n` and `m` are the globals of util.h, as in the reference.
#ifndef SAMPLE_H
#define SAMPLE_H
#include "util.h"

inline CSRSubGraph fullGraph(int *ptr, int *idx) { return CSRSubGraph(NULL, ptr, idx, n, m); }

namespace gnnagg_compat {
inline CSRSubGraph sample(int *active_vertex, int *ptr, int *idx, int fanout, int layer_num)
{
    gnnagg_aggregator *walker = NULL;
    checkGnnagg(gnnagg_create(ptr, idx, NULL, NULL, n, m, &walker));
    int *vertexset = NULL, *sub_ptr = NULL, *sub_idx = NULL, sub_v = 0, sub_e = 0;
    // seed 123: the value the reference feeds initRNG (:282)
    checkGnnagg(gnnagg_sample_subgraph(walker, active_vertex, fanout, layer_num, 123, &vertexset, &sub_ptr, &sub_idx, &sub_v,
                                       &sub_e, NULL));
    gnnagg_destroy(walker);
    dbg(sub_v);
    dbg(sub_e);
    return CSRSubGraph(vertexset, sub_ptr, sub_idx, sub_v, sub_e);  // arrays owned by the caller / its aggregator
}
}  // namespace gnnagg_compat

// active_vertex [n] holds 0/1 seed flags and is updated to the set after layer_num - 1 hops
inline CSRSubGraph sampleVertex(int *&active_vertex, int *ptr, int *idx, int layer_num = 1)
{
    return gnnagg_compat::sample(active_vertex, ptr, idx, 0, layer_num);
}

// at most neighbor_num neighbours per row (fixed-fanout stratified sampling; see gnnagg.h for how this re-specifies
// the reference's broken mark bookkeeping)
inline CSRSubGraph sampleVertexSampleNeighbor(int *&active_vertex, int *ptr, int *idx, int neighbor_num, int layer_num = 1)
{
    return gnnagg_compat::sample(active_vertex, ptr, idx, neighbor_num, layer_num);
}
#endif
