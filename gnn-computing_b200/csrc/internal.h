// internal.h -- private declarations shared by the translation units of libgnnagg.so
#pragma once
#include <stdint.h>

#include <vector>

struct gnnagg_schedule {
    int kind = 3;
    bool has_val = false;
    bool want_perm = false;          // also record perm[k] = CSR edge id stored at scheduled position k
    std::vector<int> ptr, idx, target, perm;
    std::vector<float> val;
};

namespace gnnagg {

// records a thread-local message and returns `code`
int set_error(int code, const char *msg);

// host schedule builder (host_prep.cpp)
int schedule_build(int kind, const int *ptr, const int *idx, const float *val, int num_v, int num_e, int par_num,
                   int neighbor_num, int total_num_v, gnnagg_schedule *s);

// dense combination on tcgen05 (dense_tc.cu); stream is a cudaStream_t
int dense_nn_launch(const float *A, const float *B, float *C, int64_t M, int N, int K, void *stream);

}  // namespace gnnagg
