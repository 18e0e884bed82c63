// Fused training-mode BatchNorm(+ReLU) for NHWC bf16 activations (fp32 statistics / parameters).
//
// Why it is in this library: in the headline workload (ResNet-18 replicas, per-worker batch 32)
// ATen's channels-last batch-norm kernels are 37 % of the device time of a parameter-server round
// (profiles/bench_log.md): PyTorch routes bf16 NHWC batch norm to its native kernels (cuDNN's fused
// NHWC path is fp16-only there), which are latency bound at ~25 us per call on 13-50 MB
// activations that stream in 2-8 us.  The kernels below are plain streaming kernels:
//
//   forward   reduce<false>  : per-CTA partial (sum, sum of squares) per channel, 16-byte loads
//             finalize<false>: one warp per channel folds the partials in fp64 into mean / invstd /
//                              running statistics / fused scale+shift
//             apply          : y = [relu](x * scale[c] + shift[c] [+ residual])
//   backward  reduce<true>   : per-CTA partial (sum dy', sum dy' * xhat), dy' = dy * [y > 0] (mask
//                              from the saved output when a residual was added, else recomputed
//                              from x)
//             finalize<true> : dgamma, dbeta (written straight into the gradient arena when the
//                              layer is in direct-gradient mode) and the per-channel coefficients
//             bwd_apply      : dx = P * dy' + Q * x + S, optionally d(residual) = dy'
// No atomics, deterministic summation order.  (A last-CTA-finalizes variant that saves the middle
// launch was measured 10x slower: one CTA walking 592 x C partials is L2-latency bound, ~70 us.)
//
// Activations are [R = N*H*W rows][C channels] with C % 8 == 0 (every ResNet width).
#include <cuda_bf16.h>

#include "bn.h"

namespace {

constexpr int kThreads = 256;

struct alignas(16) Bf8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack(const Bf8& p, float (&f)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(p.v[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ Bf8 pack(const float (&f)[8]) {
  Bf8 p;
#pragma unroll
  for (int k = 0; k < 4; ++k) p.v[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
  return p;
}

// One warp per channel: lanes stride over the per-CTA partials, shuffle-reduce in fp64.
__device__ __forceinline__ void reduce_channel(const float* partial, int nblocks, int C, int c, double& s,
                                               double& q) {
  const int lane = threadIdx.x & 31;
  double ls = 0.0, lq = 0.0;
#pragma unroll 4
  for (int b = lane; b < nblocks; b += 32) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)b * C + c) * 2));
    ls += (double)t.x;
    lq += (double)t.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ls += __shfl_xor_sync(0xffffffffu, ls, o);
    lq += __shfl_xor_sync(0xffffffffu, lq, o);
  }
  s = ls;
  q = lq;
}

// Per-channel epilogues of the two reductions.
struct Finalize {
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
  float* dgamma;
  float* dbeta;
  float* coef;   // [3][C]:  dx = P * dy' + Q * x + S
  long long* num_batches_tracked;  // incremented once per training forward (may be null)
  float eps, momentum;

  __device__ __forceinline__ void forward(int c, long long R, double s, double q) const {
    const double m = s / (double)R;
    double var = q / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)m;
    invstd[c] = is;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * is;
    shift[c] = b - (float)m * g * is;
    if (running_mean) {
      const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  }
  // P = gamma*invstd,  Q = -P*c2*invstd,  S = -P*c1 + P*c2*invstd*mean
  // (c1 = sum dy'/R, c2 = sum dy' xhat / R)
  __device__ __forceinline__ void backward(int c, int C, long long R, double s, double q) const {
    if (dbeta) dbeta[c] = (float)s;
    if (dgamma) dgamma[c] = (float)q;
    const double g = gamma ? (double)gamma[c] : 1.0;
    const double is = (double)invstd[c], mu = (double)mean[c];
    const double P = g * is, c1 = s / (double)R, c2 = q / (double)R;
    coef[c] = (float)P;
    coef[C + c] = (float)(-P * c2 * is);
    coef[2 * C + c] = (float)(-P * c1 + P * c2 * is * mu);
  }
};

// Thread layout shared by the two reduction kernels: cg = C / 8 channel groups along x,
// rows_per_iter = kThreads / cg row lanes; each CTA owns a contiguous slab of rows.
template <bool BWD>
__global__ void __launch_bounds__(kThreads) reduce_partial_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, long long R, int C,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, int relu, const __nv_bfloat16* __restrict__ ymask,
    float* __restrict__ partial) {
  extern __shared__ float red[];  // [rows_per_iter][C][2]
  const int cg = C >> 3;
  const int lanes = kThreads / cg;
  const int g = threadIdx.x % cg, rl = threadIdx.x / cg;
  const long long per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  const long long r1 = (r0 + per < R) ? r0 + per : R;
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a0[k] = a1[k] = 0.f;
  float sc[8], sh[8], mu[8], is[8];
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = scale[g * 8 + k];
      sh[k] = shift[g * 8 + k];
      mu[k] = mean[g * 8 + k];
      is[k] = invstd[g * 8 + k];
    }
  }
  if (rl < lanes) {
#pragma unroll 8
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const Bf8 px = *reinterpret_cast<const Bf8*>(x + r * C + g * 8);
      float xf[8];
      unpack(px, xf);
      if (!BWD) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a0[k] += xf[k];
          a1[k] = fmaf(xf[k], xf[k], a1[k]);
        }
      } else {
        const Bf8 pd = *reinterpret_cast<const Bf8*>(dy + r * C + g * 8);
        float df[8];
        unpack(pd, df);
        if (ymask != nullptr) {
          float yf[8];
          unpack(*reinterpret_cast<const Bf8*>(ymask + r * C + g * 8), yf);
#pragma unroll
          for (int k = 0; k < 8; ++k) df[k] = yf[k] > 0.f ? df[k] : 0.f;
        } else if (relu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) df[k] = fmaf(xf[k], sc[k], sh[k]) > 0.f ? df[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a0[k] += df[k];
          a1[k] = fmaf(df[k], (xf[k] - mu[k]) * is[k], a1[k]);
        }
      }
    }
  }
  // reduce across row lanes through shared memory
  float* mine = red + ((size_t)rl * C + g * 8) * 2;
  if (rl < lanes) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mine[2 * k] = a0[k];
      mine[2 * k + 1] = a1[k];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < C * 2; t += kThreads) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(size_t)l * C * 2 + t];
    partial[(size_t)blockIdx.x * C * 2 + t] = s;
  }
}

// One warp per channel; partial loads of a lane are independent (unrolled), the fold is fp64.
template <bool BWD>
__global__ void __launch_bounds__(kThreads) finalize_kernel(const float* __restrict__ partial, int nblocks,
                                                           int C, long long R, const Finalize fin) {
  const int c = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  reduce_channel(partial, nblocks, C, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  if (BWD) fin.backward(c, C, R, s, q);
  else fin.forward(c, R, s, q);
}

// The grid-stride (gridDim * 256) is a multiple of cg whenever cg divides 256 (every power-of-two
// width), so a thread always sees the same channel group and keeps its constants in registers.
__global__ void __launch_bounds__(kThreads) apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ y, long long total8,
                                                        int C, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int relu,
                                                        const __nv_bfloat16* __restrict__ res) {
  const int cg = C >> 3;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long u0 = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool fixed = (kThreads % cg) == 0;
  float sc[8], sh[8];
  if (fixed) {
    const int g = (int)(u0 % cg);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = scale[g * 8 + k];
      sh[k] = shift[g * 8 + k];
    }
  }
  for (long long u = u0; u < total8; u += stride) {
    if (!fixed) {
      const int g = (int)(u % cg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        sc[k] = __ldg(scale + g * 8 + k);
        sh[k] = __ldg(shift + g * 8 + k);
      }
    }
    float f[8], rf[8];
    unpack(reinterpret_cast<const Bf8*>(x)[u], f);
    if (res != nullptr) {
      unpack(reinterpret_cast<const Bf8*>(res)[u], rf);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) rf[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = fmaf(f[k], sc[k], sh[k]) + rf[k];
      f[k] = (relu && v < 0.f) ? 0.f : v;
    }
    reinterpret_cast<Bf8*>(y)[u] = pack(f);
  }
}

__global__ void __launch_bounds__(kThreads) bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
    __nv_bfloat16* __restrict__ dx, long long total8, int C, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ coef, int relu,
    const __nv_bfloat16* __restrict__ ymask, __nv_bfloat16* __restrict__ dres) {
  const int cg = C >> 3;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long u0 = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool fixed = (kThreads % cg) == 0;
  float sc[8], sh[8], P[8], Q[8], S[8];
  auto load_consts = [&](int g) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      sc[k] = __ldg(scale + c);
      sh[k] = __ldg(shift + c);
      P[k] = __ldg(coef + c);
      Q[k] = __ldg(coef + C + c);
      S[k] = __ldg(coef + 2 * C + c);
    }
  };
  if (fixed) load_consts((int)(u0 % cg));
  for (long long u = u0; u < total8; u += stride) {
    if (!fixed) load_consts((int)(u % cg));
    float xf[8], df[8];
    unpack(reinterpret_cast<const Bf8*>(x)[u], xf);
    unpack(reinterpret_cast<const Bf8*>(dy)[u], df);
    if (ymask != nullptr) {
      float yf[8];
      unpack(reinterpret_cast<const Bf8*>(ymask)[u], yf);
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = yf[k] > 0.f ? df[k] : 0.f;
    } else if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = fmaf(xf[k], sc[k], sh[k]) > 0.f ? df[k] : 0.f;
    }
    if (dres != nullptr) reinterpret_cast<Bf8*>(dres)[u] = pack(df);
#pragma unroll
    for (int k = 0; k < 8; ++k) df[k] = fmaf(P[k], df[k], fmaf(Q[k], xf[k], S[k]));
    reinterpret_cast<Bf8*>(dx)[u] = pack(df);
  }
}

int reduce_blocks(long long R, int sm_count) {
  // two resident CTAs per SM saturate HBM with the 8-deep unrolled 16-byte loads; more CTAs only
  // lengthen the finalize
  long long b = (long long)sm_count * 2;
  if (b > R / 64) b = R / 64;
  if (b < 1) b = 1;
  return (int)b;
}

int stream_blocks(long long total8, int sm_count) {
  long long b = (total8 + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int bz_bn_partial_blocks(long long R, int sm_count) { return reduce_blocks(R, sm_count); }

static int check_shape(long long R, int C) {
  if (C < 8 || (C % 8) != 0 || C > 4096 || R < 1) return (int)cudaErrorInvalidValue;
  if (kThreads / (C >> 3) < 1) return (int)cudaErrorInvalidValue;  // C <= 2048
  return 0;
}

static Finalize make_finalize(const BzBnArgs* a) {
  Finalize f;
  f.gamma = a->gamma;
  f.beta = a->beta;
  f.running_mean = a->running_mean;
  f.running_var = a->running_var;
  f.mean = a->mean;
  f.invstd = a->invstd;
  f.scale = a->scale;
  f.shift = a->shift;
  f.dgamma = a->dgamma;
  f.dbeta = a->dbeta;
  f.coef = a->coef;
  f.num_batches_tracked = a->num_batches_tracked;
  f.eps = a->eps;
  f.momentum = a->momentum;
  return f;
}

int bz_bn_forward(const BzBnArgs* a, int sm_count, cudaStream_t stream) {
  if (int e = check_shape(a->R, a->C)) return e;
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  auto* y = reinterpret_cast<__nv_bfloat16*>(a->y);
  const int C = a->C;
  if (a->training) {
    const int nb = reduce_blocks(a->R, sm_count);
    const int lanes = kThreads / (C >> 3);
    const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
    reduce_partial_kernel<false><<<nb, kThreads, smem, stream>>>(x, nullptr, a->R, C, nullptr, nullptr, nullptr,
                                                                nullptr, 0, nullptr, a->partial);
    finalize_kernel<false><<<(C + 7) / 8, kThreads, 0, stream>>>(a->partial, nb, C, a->R, make_finalize(a));
  }
  const long long total8 = a->R * (C >> 3);
  apply_kernel<<<stream_blocks(total8, sm_count), kThreads, 0, stream>>>(
      x, y, total8, C, a->scale, a->shift, a->relu, reinterpret_cast<const __nv_bfloat16*>(a->res));
  return (int)cudaGetLastError();
}

int bz_bn_backward(const BzBnArgs* a, int sm_count, cudaStream_t stream) {
  if (int e = check_shape(a->R, a->C)) return e;
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  const auto* dy = reinterpret_cast<const __nv_bfloat16*>(a->dy);
  const auto* ymask = reinterpret_cast<const __nv_bfloat16*>(a->ymask);
  auto* dx = reinterpret_cast<__nv_bfloat16*>(a->dx);
  const int C = a->C;
  const int nb = reduce_blocks(a->R, sm_count);
  const int lanes = kThreads / (C >> 3);
  const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
  reduce_partial_kernel<true><<<nb, kThreads, smem, stream>>>(x, dy, a->R, C, a->scale, a->shift, a->mean,
                                                             a->invstd, a->relu, ymask, a->partial);
  finalize_kernel<true><<<(C + 7) / 8, kThreads, 0, stream>>>(a->partial, nb, C, a->R, make_finalize(a));
  const long long total8 = a->R * (C >> 3);
  bwd_apply_kernel<<<stream_blocks(total8, sm_count), kThreads, 0, stream>>>(
      x, dy, dx, total8, C, a->scale, a->shift, a->coef, a->relu, ymask,
      reinterpret_cast<__nv_bfloat16*>(a->dres));
  return (int)cudaGetLastError();
}
