// Host-callable launch API of the byzpy_b200 sm_100a kernel library.
// Every function enqueues work on `stream` and returns a cudaError_t as int
// (0 == success).  No function synchronises the device.
#pragma once
#ifndef BZ_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "common.cuh"

enum BzCwMode { BZ_CW_MEDIAN = 0, BZ_CW_TRMEAN = 1, BZ_CW_MEAMED = 2, BZ_CW_MEAN = 3 };

struct BzCwArgs {
  RowTable rows;      // n real rows
  ScaleTable scales;  // per-row scale folded into the load
  int n;
  VirtRows virt;      // synthesised rows appended after the real ones
  int f;              // trim / meamed parameter
  int mode;           // BzCwMode
  long long off;      // first coordinate
  long long len;      // number of coordinates
  float* out;         // aggregated vector (may be nullptr when only updating)
  UpdTable upd;       // optional fused optimizer step
  int impl;           // 0 = auto, 1 = direct register loads, 2 = cp.async-staged pipeline,
                      // 3 = warp-tiled pipeline (many rows; falls back to auto where it does not apply)
};

// Coordinate-wise family (median / trimmed mean / mean-of-medians / mean).
int bz_cw_select(const BzCwArgs* args, int sm_count, cudaStream_t stream);

// Y[r, :] = sum_i W[r, i] * scale_i * X_i[:]   (m <= 8 output rows, W on device)
struct BzWsumArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  int m;
  const float* W;  // (m, n) row-major, device memory
  long long off, len;
  float* out[8];   // m output row pointers
  UpdTable upd;    // optional fused optimizer step using output row 0
};
int bz_wsum(const BzWsumArgs* args, int sm_count, cudaStream_t stream);

// One-pass form for 8 < m <= 128 output rows (register-tiled fp32 GEMM over shared-memory tiles);
// `len` must be a multiple of bz_wsum_multi_tile() coordinates, rows / outputs 16-byte aligned.
struct BzWsumMultiArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  int m;
  const float* W;   // (m, n) row-major, device memory
  long long off, len;
  struct { float* p[BZ_MAXN]; } out;   // m output row pointers
};
int bz_wsum_multi_tile();
int bz_wsum_multi(const BzWsumMultiArgs* args, int sm_count, cudaStream_t stream);

// Gram matrix G = (S X)(S X)^T for n <= 128 rows, fp32 CUDA-core path
// (deterministic two-stage split-K).  `partials` is scratch of
// gram_partial_elems(n, sm_count) floats.  G is (n, n) row-major.
struct BzGramArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  long long off, len;
  float* partials;
  int num_partials;
  double* G64;  // optional fp64 output (may be nullptr)
  float* G;     // fp32 output
  // Optional fused auxiliary row (n <= 16): the coordinate-wise lower median of the scaled rows is
  // computed while the rows stream by, written to aux_median (len floats, same offset as the rows)
  // and appended as row n of the Gram matrix, which is then (n+1) x (n+1).  This is the start point of
  // Weiszfeld / the centre of centred clipping: fusing it saves a full pass over the n x d matrix.
  float* aux_median;
};
int bz_gram_partials_needed(int n, int sm_count);
int bz_gram(const BzGramArgs* args, int sm_count, cudaStream_t stream);

// Column statistics: out = a*mean + b*std (population) over n rows.
struct BzColStatArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  float a, b;
  long long off, len;
  float* out;
};
int bz_colstat(const BzColStatArgs* args, int sm_count, cudaStream_t stream);

// Element-wise helpers (attacks / flat optimizer).
int bz_scale_copy(const float* src, float* dst, float scale, long long len, int sm_count,
                  cudaStream_t stream);
int bz_fill(float* dst, float value, long long len, int sm_count, cudaStream_t stream);
int bz_gaussian(float* dst, float mu, float sigma, unsigned long long seed,
                unsigned long long offset, long long len, int sm_count, cudaStream_t stream);
int bz_sgd(const float* grad, const UpdTable* upd, long long len, int sm_count,
           cudaStream_t stream);

// uint8 [..., C] -> bf16 (x - mean[c]) * scale[c], same element order (C <= 8, 16-byte aligned buffers).
int bz_u8_affine(const void* in, void* out, long long n, int C, const float* mean, const float* scale,
                 int sm_count, cudaStream_t stream);
