// dense.h -- compatibility layer: matmul_NN of the reference (include/dense.h:4-23, cublasSgemm +
// cublasSgeam transpose) as the library's tcgen05 3xTF32 product.  C[inM,inN] = A[inM,inK]*B[inK,inN],
// all row-major; `tmp` (the reference's transpose scratch) is unused.
#ifndef DENSE_H
#define DENSE_H
#include "util.h"

inline void matmul_NN(float *A, float *B, float *C, int inM, int inN, int inK, float *tmp)
{
    (void)tmp;
    checkGnnagg(gnnagg_dense_nn(A, B, C, inM, inN, inK, NULL));
}
#endif
