// sample.h -- compatibility layer: only fullGraph() of the reference's include/sample.h:126-129 is on
// the aggregation path (the GPU samplers there are called by no driver).
#ifndef SAMPLE_H
#define SAMPLE_H
#include "util.h"

inline CSRSubGraph fullGraph(int *ptr, int *idx) { return CSRSubGraph(NULL, ptr, idx, n, m); }
#endif
