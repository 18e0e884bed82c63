// Fused NHWC bf16 BatchNorm(+ReLU) forward / backward launchers (bn.cu).
#pragma once
#include <cuda_runtime.h>

struct BzBnArgs {
  const void* x;          // bf16 [R][C]
  void* y;                // bf16 [R][C]            (forward)
  const void* dy;         // bf16 [R][C]            (backward)
  void* dx;               // bf16 [R][C]            (backward)
  long long R;
  int C;
  const float* gamma;     // may be null (affine=False)
  const float* beta;
  float* running_mean;    // may be null
  float* running_var;
  float* mean;            // [C] batch mean            (written in training forward, read in backward)
  float* invstd;          // [C]
  float* scale;           // [C] gamma * invstd        (eval mode: precomputed by the caller)
  float* shift;           // [C] beta - mean * scale
  float* partial;         // [bz_bn_partial_blocks][C][2] scratch
  float* dgamma;          // [C] (backward, may be null)
  float* dbeta;
  float* coef;            // [3][C] scratch (backward)
  const void* res;        // bf16 [R][C] residual added before the ReLU (forward, may be null)
  const void* ymask;      // bf16 [R][C] forward output; ReLU mask source in backward (null: recompute)
  void* dres;             // bf16 [R][C] gradient of the residual = masked dy (backward, may be null)
  long long* num_batches_tracked;  // int64 scalar bumped by the training forward (may be null)
  int* launches;          // host out-param: kernels launched by this call (may be null)
  float eps, momentum;
  int relu;
  int training;
};

int bz_bn_partial_blocks(long long R, int sm_count);
int bz_bn_forward(const BzBnArgs* args, int sm_count, cudaStream_t stream);
int bz_bn_backward(const BzBnArgs* args, int sm_count, cudaStream_t stream);
