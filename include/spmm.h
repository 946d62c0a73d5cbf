// spmm.h -- compatibility layer for include/spmm.h of the reference: the naive thread-per-row
// spmm<LENFEATURE> kernel and the result validators valid() / validReordered().
#ifndef SPMM_H
#define SPMM_H
#include "util.h"

#define TB 128

// Naive SpMM, one thread per row, launched by the caller as spmm<F><<<ceil(numV/TB), TB>>>(...)
// (spmm.h:223-265).  Rows without neighbours are left untouched, as in the reference.
template <int LENFEATURE>
__global__ void spmm(int numV, int *ptr, int *idx, float *val, float *denseInput, float *denseOutput)
{
    static_assert(LENFEATURE % 4 == 0, "feature length must be a multiple of 4");
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= numV) return;
    const int first = ptr[row], last = ptr[row + 1];
    if (first == last) return;
    float4 acc[LENFEATURE / 4];
#pragma unroll
    for (int q = 0; q < LENFEATURE / 4; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = first; e < last; ++e) {
        const float w = val[e];
        const float4 *x = reinterpret_cast<const float4 *>(denseInput + (size_t)idx[e] * LENFEATURE);
#pragma unroll
        for (int q = 0; q < LENFEATURE / 4; ++q) {
            const float4 v = x[q];
            acc[q].x += w * v.x, acc[q].y += w * v.y, acc[q].z += w * v.z, acc[q].w += w * v.w;
        }
    }
    float4 *y = reinterpret_cast<float4 *>(denseOutput + (size_t)row * LENFEATURE);
#pragma unroll
    for (int q = 0; q < LENFEATURE / 4; ++q) y[q] = acc[q];
}

// number of elements whose relative error exceeds 1e-2 (spmm.h:11-21,35-69)
inline int valid(float *y, float *y2, int num)
{
    int diff = 33;
    checkGnnagg(gnnagg_validate(y, y2, num, &diff, NULL));
    return diff;
}

// same through the reorder permutation `rows` published by load_graph (spmm.h:23-33,71-91)
inline int validReordered(float *y, float *y2, int num_v, int feature_len)
{
    if (rows == NULL) return valid(y, y2, num_v * feature_len);
    int *d_map = NULL, diff = 33;
    checkCudaErrors(cudaMalloc2((void **)&d_map, num_v * sizeof(int)));
    checkCudaErrors(cudaMemcpy(d_map, rows, num_v * sizeof(int), cudaMemcpyHostToDevice));
    checkGnnagg(gnnagg_validate_reordered(y, y2, d_map, num_v, feature_len, &diff, NULL));
    cudaFree(d_map);
    return diff;
}
#endif
