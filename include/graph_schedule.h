// graph_schedule.h -- compatibility layer: the three host schedule functions of the reference
// (include/graph_schedule.h:17,91,156) with their exact parameter lists, implemented on
// gnnagg_schedule_build of libgnnagg.so (count / scan / fill passes, one sweep over the edges;
// bit-exact outputs, see tests/test_host_prep.py).
#ifndef GRAPH_SCHEDULE_H
#define GRAPH_SCHEDULE_H
#include <vector>

#include "util.h"

enum Schedule { locality, neighbor_grouping, locality_neighbor_grouping, nop };

namespace gnnagg_compat {
// appends a schedule to the caller's vectors the way the reference's push_back loops do
// (ptr_vec gains a leading 0 followed by the group ends)
inline void emit(gnnagg_schedule *s, std::vector<int> *ptr_vec, std::vector<int> *idx_vec,
                 std::vector<int> *target_vec, std::vector<float> *val_vec)
{
    const int64_t g = gnnagg_schedule_num_target(s), e = gnnagg_schedule_num_edges(s);
    const int *p = gnnagg_schedule_ptr(s), *t = gnnagg_schedule_target(s), *i = gnnagg_schedule_idx(s);
    ptr_vec->insert(ptr_vec->end(), p, p + g + 1);
    target_vec->insert(target_vec->end(), t, t + g);
    idx_vec->insert(idx_vec->end(), i, i + e);
    const float *v = gnnagg_schedule_val(s);
    if (val_vec && v) val_vec->insert(val_vec->end(), v, v + e);
    gnnagg_schedule_free(s);
}
inline int edge_count(const int *ptr, int num_v) { return num_v > 0 ? ptr[num_v] : 0; }
}  // namespace gnnagg_compat

inline void locality_schedule(int *ptr, int *idx, int par_num, int num_v, std::vector<int> *ptr_vec,
                              std::vector<int> *idx_vec, std::vector<int> *target_vec, int total_num_v,
                              float *val = NULL, std::vector<float> *val_vec = NULL)
{
    timestamp(t0);
    gnnagg_schedule *s = NULL;
    checkGnnagg(gnnagg_schedule_build(GNNAGG_SCHED_LOCALITY, ptr, idx, val, num_v, gnnagg_compat::edge_count(ptr, num_v),
                                      par_num, 0, total_num_v, &s));
    gnnagg_compat::emit(s, ptr_vec, idx_vec, target_vec, val_vec);
    timestamp(t1);
    double locality_schedule_time = getDuration(t0, t1);
    dbg(locality_schedule_time);
}

inline void neighbor_grouping_schedule(int *ptr, int *idx, int neighbor_num, int num_v, int num_e,
                                       std::vector<int> *ptr_vec, std::vector<int> *idx_vec,
                                       std::vector<int> *target_vec)
{
    assert(ptr != NULL);
    assert(idx != NULL);
    timestamp(t0);
    gnnagg_schedule *s = NULL;
    checkGnnagg(gnnagg_schedule_build(GNNAGG_SCHED_NEIGHBOR_GROUPING, ptr, idx, NULL, num_v, num_e, 0, neighbor_num,
                                      num_v, &s));
    gnnagg_compat::emit(s, ptr_vec, idx_vec, target_vec, NULL);
    dbg(target_vec->size());
    dbg(num_e);
    timestamp(t1);
    double neighbor_grouping_schedule_time = getDuration(t0, t1);
    dbg(neighbor_grouping_schedule_time);
}

inline void localityNeighborGrouping(int *ptr, int *idx, int par_num, int neighbor_num, int num_v,
                                     std::vector<int> *ptr_vec, std::vector<int> *idx_vec,
                                     std::vector<int> *target_vec, int total_num_v, float *val = NULL,
                                     std::vector<float> *val_vec = NULL)
{
    timestamp(t0);
    gnnagg_schedule *s = NULL;
    checkGnnagg(gnnagg_schedule_build(GNNAGG_SCHED_LOCALITY_NEIGHBOR_GROUPING, ptr, idx, val, num_v,
                                      gnnagg_compat::edge_count(ptr, num_v), par_num, neighbor_num, total_num_v, &s));
    gnnagg_compat::emit(s, ptr_vec, idx_vec, target_vec, val_vec);
    timestamp(t1);
    dbg(getDuration(t0, t1));
}
#endif
