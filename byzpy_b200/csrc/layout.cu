// Conv-weight layout casts between the fp32 OIHW master/gradient arena and cuDNN's bf16
// channels-last (KRSC) operands.
//
//   to_shadow : fp32 [K][C][RS]  ->  bf16 [K][RS][C]     (weights, before the forward pass)
//   to_grad   : bf16 [K][RS][C]  ->  fp32 [K][C][RS]     (weight gradients, after wgrad)
//
// ATen does these as a strided `copy_` whose reads are uncoalesced: 46 us for a 512x512x3x3 filter
// (profiles/bench_log.md), three of those per replica step.  Here a CTA stages a few filters in shared
// memory, so both the global read and the global write are contiguous; one launch handles up to 64
// tensors (a whole ResNet-50's convolutions) from a by-value table in the kernel parameters.
#include <cuda_bf16.h>
#include <stdint.h>

#include "layout.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxElems = 9216;  // per-CTA staging buffer (fp32), 36 KB (+ padding below)

// Filters per CTA so that a CTA moves a few thousand elements.
__host__ __device__ inline int filters_per_cta(int C, int RS) {
  const int per = C * RS;
  int k = kMaxElems / (per + RS);   // + RS: room for the padded layout of to_grad
  if (k < 1) k = 1;
  if (k > 64) k = 64;
  return k;
}

template <bool TO_GRAD>
__global__ void __launch_bounds__(kThreads) krsc_cast_kernel(const __grid_constant__ BzCastTable t) {
  extern __shared__ float buf[];
  // locate this CTA's tensor
  int e = 0;
#pragma unroll 1
  while (e + 1 < t.count && (int)blockIdx.x >= t.e[e + 1].cta_start) ++e;
  const BzCastEntry& en = t.e[e];
  const int C = en.C, RS = en.RS, K = en.K;
  const int kpb = filters_per_cta(C, RS);
  const int k0 = ((int)blockIdx.x - en.cta_start) * kpb;
  const int nk = (K - k0 < kpb) ? K - k0 : kpb;
  if (nk <= 0) return;
  const int per = C * RS;
  const int n = nk * per;
  if (TO_GRAD) {
    // src bf16 [k][rs][c] -> smem [(k*RS + rs) * (C + 1) + c]  (odd pitch: conflict-light transposed reads)
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(en.src) + (size_t)k0 * per;
    float* dst = reinterpret_cast<float*>(en.dst) + (size_t)k0 * per;
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const int row = i / C, c = i - row * C;
      buf[row * (C + 1) + c] = __bfloat162float(src[i]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += kThreads) {
      const int kk = j / per, r = j - kk * per;
      const int c = r / RS, rs = r - c * RS;
      dst[j] = buf[(kk * RS + rs) * (C + 1) + c];
    }
  } else {
    // src fp32 [k][c][rs] -> smem linear; read with stride RS (odd for 3x3 / 7x7 -> conflict free)
    const float* src = reinterpret_cast<const float*>(en.src) + (size_t)k0 * per;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(en.dst) + (size_t)k0 * per;
    for (int i = threadIdx.x; i < n; i += kThreads) buf[i] = src[i];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += kThreads) {
      const int kk = j / per, r = j - kk * per;
      const int rs = r / C, c = r - rs * C;
      dst[j] = __float2bfloat16_rn(buf[kk * per + c * RS + rs]);
    }
  }
}

}  // namespace

int bz_krsc_cast_ctas(int K, int C, int RS) {
  const int kpb = filters_per_cta(C, RS);
  return (K + kpb - 1) / kpb;
}

int bz_krsc_cast(const BzCastTable* table, int to_grad, cudaStream_t stream) {
  const BzCastTable& t = *table;
  if (t.count < 1 || t.count > BZ_CAST_MAX) return (int)cudaErrorInvalidValue;
  int total = 0;
  size_t smem = 0;
  for (int i = 0; i < t.count; ++i) {
    const BzCastEntry& e = t.e[i];
    if (e.K < 1 || e.C < 1 || e.RS < 1 || e.cta_start != total) return (int)cudaErrorInvalidValue;
    const int kpb = filters_per_cta(e.C, e.RS);
    const size_t need = (size_t)kpb * e.RS * (e.C + 1) * sizeof(float);
    if (need > 200 * 1024) return (int)cudaErrorInvalidValue;  // C*RS > ~50k elements per filter
    if (need > smem) smem = need;
    total += bz_krsc_cast_ctas(e.K, e.C, e.RS);
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(krsc_cast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024);
    cudaError_t e2 = cudaFuncSetAttribute(krsc_cast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024);
    if (e1 != cudaSuccess) return (int)e1;
    if (e2 != cudaSuccess) return (int)e2;
    configured = true;
  }
  if (to_grad) krsc_cast_kernel<true><<<total, kThreads, smem, stream>>>(t);
  else krsc_cast_kernel<false><<<total, kThreads, smem, stream>>>(t);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Space-to-depth form of the ResNet stem (7x7 / stride 2 / pad 3 convolution over 3 channels).
//
// cuDNN runs the 3-channel stem with a legacy sm_80 kernel (198 us forward + 122 us weight gradient
// per batch-32 replica, 13 % of a replica step; padding the input to 4 or 8 channels is not faster --
// bench/conv_stem.py).  A 2x2 space-to-depth of the zero-padded input turns it into a dense 4x4 /
// stride 1 convolution over 12 (stored as 16) channels, which maps onto the sm_100 tensor-core
// kernels:
//     x'[n, bi, bj, (p*2+q)*3 + c] = xpad[n, 2*bi + p, 2*bj + q, c],   xpad = x shifted by the pad of 3
//     w'[k, dr, ds, (p*2+q)*3 + c] = w[k, c, 2*dr + p, 2*ds + q]        (0 where the index would be 7)
// These kernels produce x' (optionally straight from the uint8 image batch, normalisation fused), w'
// (bf16, channels-last) and scatter the gradient of w' back into the fp32 OIHW gradient of w.
namespace {

__device__ __forceinline__ uint4 pack8(const float* f) {     // 8 floats -> 8 bf16 in one 128-bit word
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename TIn>
__device__ __forceinline__ float load_px(const TIn* p);
template <>
__device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) {
  return (float)*p;
}
template <>
__device__ __forceinline__ float load_px<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// one thread per (n, bi, bj): 16 output channels = 32 bytes
template <typename TIn>
__global__ void __launch_bounds__(256) s2d_pack_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                      int N, int H, int W, int Hb, int Wb, float m0, float m1,
                                                      float m2, float s0, float s1, float s2) {
  const long long total = (long long)N * Hb * Wb;
  const long long u = (long long)blockIdx.x * 256 + threadIdx.x;
  if (u >= total) return;
  const int bj = (int)(u % Wb);
  const long long t = u / Wb;
  const int bi = (int)(t % Hb);
  const int n = (int)(t / Hb);
  const float mean[3] = {m0, m1, m2}, scale[3] = {s0, s1, s2};
  float f[16];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * bi + p - 3, j = 2 * bj + q - 3;
      const bool in = (i >= 0 && i < H && j >= 0 && j < W);
      const TIn* src = x + (((long long)n * H + (in ? i : 0)) * W + (in ? j : 0)) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[(p * 2 + q) * 3 + c] = in ? (load_px<TIn>(src + c) - mean[c]) * scale[c] : 0.f;
    }
  }
#pragma unroll
  for (int k = 12; k < 16; ++k) f[k] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(out + u * 16);
  dst[0] = pack8(f);
  dst[1] = pack8(f + 8);
}

// w fp32 [K][3][7][7] -> w' bf16 [K][4][4][16]
__global__ void stem_weight_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * 256) return;
  const int ch = idx & 15, ds = (idx >> 4) & 3, dr = (idx >> 6) & 3, k = idx >> 8;
  float v = 0.f;
  if (ch < 12) {
    const int c = ch % 3, pq = ch / 3, p = pq >> 1, q = pq & 1;
    const int r = 2 * dr + p, s = 2 * ds + q;
    if (r < 7 && s < 7) v = w[((k * 3 + c) * 7 + r) * 7 + s];
  }
  wp[idx] = __float2bfloat16_rn(v);
}

// dW' bf16 [K][4][4][16] -> dW fp32 [K][3][7][7]
__global__ void stem_grad_unpack_kernel(const __nv_bfloat16* __restrict__ gp, float* __restrict__ g, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * 147) return;
  const int s = idx % 7, r = (idx / 7) % 7, c = (idx / 49) % 3, k = idx / 147;
  const int dr = r >> 1, p = r & 1, ds = s >> 1, q = s & 1;
  g[idx] = __bfloat162float(gp[((k * 4 + dr) * 4 + ds) * 16 + (p * 2 + q) * 3 + c]);
}

}  // namespace

int bz_s2d_pack(const void* x, int x_is_u8, void* out, int N, int H, int W, const float* mean,
                const float* scale, cudaStream_t stream) {
  if (N < 1 || H < 1 || W < 1) return (int)cudaErrorInvalidValue;
  const int Hb = (H - 1) / 2 + 1 + 3, Wb = (W - 1) / 2 + 1 + 3;
  const long long total = (long long)N * Hb * Wb;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (x_is_u8)
    s2d_pack_kernel<uint8_t><<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(x), o, N, H, W, Hb, Wb,
                                                        mean[0], mean[1], mean[2], scale[0], scale[1], scale[2]);
  else
    s2d_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), o, N, H,
                                                              W, Hb, Wb, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
  return (int)cudaGetLastError();
}

int bz_stem_weight_pack(const float* w, void* wp, int K, cudaStream_t stream) {
  stem_weight_pack_kernel<<<(K * 256 + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wp), K);
  return (int)cudaGetLastError();
}

int bz_stem_grad_unpack(const void* gp, float* g, int K, cudaStream_t stream) {
  stem_grad_unpack_kernel<<<(K * 147 + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(gp), g,
                                                                    K);
  return (int)cudaGetLastError();
}
