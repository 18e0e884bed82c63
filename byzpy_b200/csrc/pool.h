// NHWC bf16 3x3/s2/p1 max pooling (pool.cu).  idx holds one byte (window position kh*3+kw) per output
// element.
#pragma once
#include <cuda_runtime.h>

int bz_maxpool3x3s2_forward(const void* x, void* y, void* idx, int N, int H, int W, int C, int sm_count,
                            cudaStream_t stream);
int bz_maxpool3x3s2_backward(const void* dy, const void* idx, void* dx, int N, int H, int W, int C,
                             int sm_count, cudaStream_t stream);
