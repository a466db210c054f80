// tc5_dense.cuh -- host interface of the tcgen05 dense-layer engine (tc5_dense.cu) used by the entry points of dense.cu.
#pragma once
#include "xdr_common.cuh"

namespace xdr {

// true when the engine is on and takes the shape / alignment (else the caller stays on the fp32 FMA kernels)
bool tc5_dense_fwd_ok(const float* X, const float* W, const float* X2, const float* W2, const float* Y, int64_t M, int N, int K);
bool tc5_dense_bwd_input_ok(const float* dZ, const float* W, const float* dX, int64_t M, int N, int K);
bool tc5_dense_bwd_weight_ok(const float* dZ, const float* X, const float* dW, int64_t M, int N, int K);

int tc5_dense_fwd(const float* X, const float* W, const float* bias, const float* X2, const float* W2, const int64_t* mask_ids,
                  int64_t mask_lt, int act, float* Y, int64_t M, int N, int K, cudaStream_t stream);
int tc5_dense_bwd_input(const float* dZ, const float* W, const int64_t* mask_ids, int64_t mask_lt, float* dX, int64_t M, int N,
                        int K, int accumulate, cudaStream_t stream);
int tc5_dense_bwd_weight(const float* dZ, const float* X, const int64_t* mask_ids, int64_t mask_lt, float* dW, float* db, int64_t M,
                         int N, int K, cudaStream_t stream);

}  // namespace xdr
