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
