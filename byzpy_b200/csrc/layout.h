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

// Space-to-depth form of the 7x7/s2/p3 3-channel stem convolution (see layout.cu).
//   x : uint8 (x_is_u8, normalised as (x - mean[c]) * scale[c]) or bf16, NHWC [N][H][W][3]
//   out: bf16 [N][Hb][Wb][16] with Hb = (H-1)/2 + 4, Wb = (W-1)/2 + 4
int bz_s2d_pack(const void* x, int x_is_u8, void* out, int N, int H, int W, const float* mean,
                const float* scale, cudaStream_t stream);
int bz_stem_weight_pack(const float* w, void* wp, int K, cudaStream_t stream);      // [K][3][7][7] -> [K][4][4][16]
int bz_stem_grad_unpack(const void* gp, float* g, int K, cudaStream_t stream);      // and back (fp32)
