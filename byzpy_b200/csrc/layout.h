// Conv-weight layout casts (layout.cu): fp32 OIHW arena <-> bf16 channels-last cuDNN operands.
#pragma once
#include <cuda_runtime.h>

#define BZ_CAST_MAX 64

struct BzCastEntry {
  const void* src;
  void* dst;
  int K, C, RS;
  int cta_start;  // first CTA of this tensor (prefix sum of bz_krsc_cast_ctas over the table)
};
struct BzCastTable {
  BzCastEntry e[BZ_CAST_MAX];
  int count;
};

int bz_krsc_cast_ctas(int K, int C, int RS);
// to_grad = 0: fp32 [K][C][RS] -> bf16 [K][RS][C];  to_grad = 1: bf16 [K][RS][C] -> fp32 [K][C][RS]
int bz_krsc_cast(const BzCastTable* table, int to_grad, cudaStream_t stream);
