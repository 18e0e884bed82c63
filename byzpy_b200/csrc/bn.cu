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
#include <stdint.h>

#include <cstdlib>
#include <utility>

#include "bn.h"

namespace {

constexpr int kThreads = 256;

// Programmatic dependent launch (sm_90+): the finalize / apply kernels of a direction are launched
// with programmaticStreamSerialization, so their CTAs are scheduled while the previous kernel of the
// chain drains; they block in griddepcontrol.wait until that kernel has completed and flushed, i.e.
// ordinary stream semantics with the ~2 us launch latency hidden.  Both instructions are no-ops in a
// kernel launched the ordinary way.  BYZPY_B200_NO_PDL=1 switches the attribute off (A/B testing).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("BYZPY_B200_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_chain(bool dependent, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                         cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (dependent && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Eight bf16 values moved as ONE 128-bit access.  (A struct of four __nv_bfloat162 is copied member by
// member -- four LDG.32 -- which capped the first version of these kernels at a quarter of the load
// width; ncu: 1.75 TB/s on the 51 MB stem activation.)
struct alignas(16) Bf8 {
  uint4 u;
};

__device__ __forceinline__ void unpack(const Bf8& p, float (&f)[8]) {
  const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // bf16 -> fp32 is a 16-bit shift
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ Bf8 pack(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  Bf8 p;
  p.u = make_uint4(w[0], w[1], w[2], w[3]);
  return p;
}
__device__ __forceinline__ Bf8 ld8(const __nv_bfloat16* p) {
  Bf8 r;
  r.u = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const Bf8& v) { *reinterpret_cast<uint4*>(p) = v.u; }

// One warp per channel: lanes stride over the per-CTA partials, shuffle-reduce in fp64.
__device__ __forceinline__ void reduce_channel(const float* partial, int nblocks, int C, int c, double& s,
                                               double& q) {
  const int lane = threadIdx.x & 31;
  double ls = 0.0, lq = 0.0;
  // batches of 8 loads per lane are issued before the first add (same reason as in the reduce kernel)
  for (int b0 = lane; b0 < nblocks; b0 += 32 * 8) {
    float2 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + 32 * j;
      t[j] = (b < nblocks) ? __ldg(reinterpret_cast<const float2*>(partial + ((size_t)b * C + c) * 2))
                           : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ls += (double)t[j].x;
      lq += (double)t[j].y;
    }
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

  // forward statistics of channel c from (sum, sum of squares); optionally stored
  __device__ __forceinline__ void forward(int c, long long R, double s, double q, float& sc, float& sh,
                                          bool store) const {
    const double m = s / (double)R;
    double var = q / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    sc = g * is;
    sh = b - (float)m * g * is;
    if (!store) return;
    mean[c] = (float)m;
    invstd[c] = is;
    scale[c] = sc;
    shift[c] = sh;
    if (running_mean) {
      const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  }
  // P = gamma*invstd,  Q = -P*c2*invstd,  S = -P*c1 + P*c2*invstd*mean
  // (c1 = sum dy'/R, c2 = sum dy' xhat / R)
  __device__ __forceinline__ void backward(int c, int C, long long R, double s, double q, float& P_, float& Q_,
                                           float& S_, bool store) const {
    const double g = gamma ? (double)gamma[c] : 1.0;
    const double is = (double)invstd[c], mu = (double)mean[c];
    const double P = g * is, c1 = s / (double)R, c2 = q / (double)R;
    P_ = (float)P;
    Q_ = (float)(-P * c2 * is);
    S_ = (float)(-P * c1 + P * c2 * is * mu);
    if (!store) return;
    if (dbeta) dbeta[c] = (float)s;
    if (dgamma) dgamma[c] = (float)q;
    if (coef) {
      coef[c] = P_;
      coef[C + c] = Q_;
      coef[2 * C + c] = S_;
    }
  }
};

// Sequential fp64 fold of the per-CTA partials of one channel (small partial counts: the
// finalize-in-prologue variants of the apply kernels, identical in every CTA).
__device__ __forceinline__ void fold_channel(const float* __restrict__ partial, int nblocks, int C, int c,
                                             double& s, double& q) {
  // all (<= 32) loads are issued before the first add: one L2 round trip instead of nblocks
  float2 t[32];
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    t[b] = (b < nblocks) ? __ldg(reinterpret_cast<const float2*>(partial + ((size_t)b * C + c) * 2))
                         : make_float2(0.f, 0.f);
  }
  double ls = 0.0, lq = 0.0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    ls += (double)t[b].x;
    lq += (double)t[b].y;
  }
  s = ls;
  q = lq;
}

// Thread layout shared by the two reduction kernels: cg = C / 8 channel groups along x,
// rows_per_iter = kThreads / cg row lanes; each CTA owns a contiguous slab of rows.
template <bool BWD>
__global__ void __launch_bounds__(kThreads) reduce_partial_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, long long R, int C,
    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
    const float* __restrict__ invstd, int relu, const __nv_bfloat16* __restrict__ ymask,
    float* __restrict__ partial) {
  extern __shared__ float red[];  // [rows_per_iter][C][2]
  pdl_launch_dependents();        // let the next kernel of the chain get scheduled behind this one
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
  // One row of this thread's channel group.
  auto consume = [&](const Bf8& px, const Bf8& pd, const Bf8& py) {
    float xf[8];
    unpack(px, xf);
    if (!BWD) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a0[k] += xf[k];
        a1[k] = fmaf(xf[k], xf[k], a1[k]);
      }
    } else {
      float df[8];
      unpack(pd, df);
      if (ymask != nullptr) {
        float yf[8];
        unpack(py, yf);
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
  };
  if (rl < lanes) {
    // Explicit load batches: U rows' 128-bit loads are issued back to back into a register array and
    // only then consumed.  (A plain `#pragma unroll 8` loop was scheduled load -> use -> load with two
    // registers' worth of loads in flight: ncu showed 1.75 TB/s and long-scoreboard stalls.)
    constexpr int U = BWD ? 4 : 8;
    long long r = r0 + rl;
    for (; r + (long long)(U - 1) * lanes < r1; r += (long long)U * lanes) {
      Bf8 bx[U], bd[BWD ? U : 1], by[BWD ? U : 1];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const long long o = (r + (long long)j * lanes) * C + g * 8;
        bx[j] = ld8(x + o);
        if (BWD) {
          bd[j] = ld8(dy + o);
          if (ymask != nullptr) by[j] = ld8(ymask + o);
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j) consume(bx[j], bd[BWD ? j : 0], by[BWD ? j : 0]);
    }
    for (; r < r1; r += lanes) {
      const long long o = r * C + g * 8;
      Bf8 pd = {}, py = {};
      if (BWD) {
        pd = ld8(dy + o);
        if (ymask != nullptr) py = ld8(ymask + o);
      }
      consume(ld8(x + o), pd, py);
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
  pdl_wait();
  pdl_launch_dependents();
  const int c = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  reduce_channel(partial, nblocks, C, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  float t0, t1, t2;
  if (BWD) fin.backward(c, C, R, s, q, t0, t1, t2, true);
  else fin.forward(c, R, s, q, t0, t1, true);
}

// The grid-stride (gridDim * 256) is a multiple of cg whenever cg divides 256 (every power-of-two
// width), so a thread always sees the same channel group and keeps its constants in registers.
//
// FIN variants (small activations): the per-channel finalize runs in the prologue of every CTA from a
// SMALL number of partials (<= 32, identical arithmetic everywhere, CTA 0 stores the statistics), which
// removes the middle launch of a direction: at 1.6-6.4 MB the three dependent launches cost 12-15 us
// while the data streams in 1-3 us (bench/bn_layers.py).
template <bool FIN>
__global__ void __launch_bounds__(kThreads) apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ y, long long total8,
                                                        int C, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int relu,
                                                        const __nv_bfloat16* __restrict__ res,
                                                        const float* __restrict__ partial, int nblocks,
                                                        long long R, const Finalize fin) {
  extern __shared__ float cs[];   // FIN: [C][2] = scale, shift
  pdl_wait();
  if (FIN) {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      double s, q;
      fold_channel(partial, nblocks, C, c, s, q);
      float sc_, sh_;
      fin.forward(c, R, s, q, sc_, sh_, blockIdx.x == 0);
      cs[2 * c] = sc_;
      cs[2 * c + 1] = sh_;
    }
    __syncthreads();
    scale = nullptr;
  }
  const int cg = C >> 3;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long u0 = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool fixed = (kThreads % cg) == 0;
  float sc[8], sh[8];
  auto load_consts = [&](int g) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      sc[k] = FIN ? cs[2 * c] : __ldg(scale + c);
      sh[k] = FIN ? cs[2 * c + 1] : __ldg(shift + c);
    }
  };
  if (fixed) load_consts((int)(u0 % cg));
  for (long long u = u0; u < total8; u += stride) {
    if (!fixed) load_consts((int)(u % cg));
    float f[8], rf[8];
    unpack(ld8(x + u * 8), f);
    if (res != nullptr) {
      unpack(ld8(res + u * 8), rf);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) rf[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = fmaf(f[k], sc[k], sh[k]) + rf[k];
      f[k] = (relu && v < 0.f) ? 0.f : v;
    }
    st8(y + u * 8, pack(f));
  }
}

template <bool FIN>
__global__ void __launch_bounds__(kThreads) bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
    __nv_bfloat16* __restrict__ dx, long long total8, int C, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ coef, int relu,
    const __nv_bfloat16* __restrict__ ymask, __nv_bfloat16* __restrict__ dres,
    const float* __restrict__ partial, int nblocks, long long R, const Finalize fin) {
  extern __shared__ float cs[];   // FIN: [C][3] = P, Q, S
  pdl_wait();
  if (FIN) {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      double s, q;
      fold_channel(partial, nblocks, C, c, s, q);
      float P_, Q_, S_;
      fin.backward(c, C, R, s, q, P_, Q_, S_, blockIdx.x == 0);
      cs[3 * c] = P_;
      cs[3 * c + 1] = Q_;
      cs[3 * c + 2] = S_;
    }
    __syncthreads();
  }
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
      P[k] = FIN ? cs[3 * c] : __ldg(coef + c);
      Q[k] = FIN ? cs[3 * c + 1] : __ldg(coef + C + c);
      S[k] = FIN ? cs[3 * c + 2] : __ldg(coef + 2 * C + c);
    }
  };
  if (fixed) load_consts((int)(u0 % cg));
  for (long long u = u0; u < total8; u += stride) {
    if (!fixed) load_consts((int)(u % cg));
    float xf[8], df[8];
    unpack(ld8(x + u * 8), xf);
    unpack(ld8(dy + u * 8), df);
    if (ymask != nullptr) {
      float yf[8];
      unpack(ld8(ymask + u * 8), yf);
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = yf[k] > 0.f ? df[k] : 0.f;
    } else if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = fmaf(xf[k], sc[k], sh[k]) > 0.f ? df[k] : 0.f;
    }
    if (dres != nullptr) st8(dres + u * 8, pack(df));
#pragma unroll
    for (int k = 0; k < 8; ++k) df[k] = fmaf(P[k], df[k], fmaf(Q[k], xf[k], S[k]));
    st8(dx + u * 8, pack(df));
  }
}

// ---------------------------------------------------------------------------------------------
// Single-launch forms for SMALL activations (<= 4 MB: the 14x14 and 7x7 stages of a ResNet at batch
// 32), built on thread-block clusters.  A cluster of kClusterSize CTAs owns a group of CGC channels
// for ALL rows: every CTA reduces its slab of rows, the per-CTA partials meet through distributed
// shared memory (mapa + ld.shared::cluster behind one barrier.cluster), every CTA finalises the
// group's statistics itself (identical fp64 fold, rank 0 stores them) and applies them to its slab,
// which is still in L2.  One launch instead of two dependent ones: at these sizes a direction costs
// ~5.5 us per dependent launch while the data streams in 1-2 us (bench/bn_layers.py).  No cross-cluster
// synchronisation exists, so nothing depends on co-residency of the whole grid.
constexpr int kClusterSize = 8;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(local), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

template <int CGC, bool BWD>
__global__ void __launch_bounds__(kThreads) bn_cluster_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out,
    long long R, int C, int relu, const __nv_bfloat16* __restrict__ res_or_mask, __nv_bfloat16* __restrict__ dres,
    const Finalize fin) {
  constexpr int NG = CGC / 8;                 // 8-channel groups per cluster
  constexpr int LANES = kThreads / NG;        // row lanes per CTA
  __shared__ float red[LANES * CGC * 2];      // 16 KB
  __shared__ float part[CGC * 2];             // this CTA's partial sums, read by the whole cluster
  __shared__ float cs[CGC * 3];               // scale, shift | P, Q, S
  const uint32_t rank = cluster_rank();
  const int c0 = (blockIdx.x / kClusterSize) * CGC;
  const int g = threadIdx.x % NG, rl = threadIdx.x / NG;
  const long long per = (R + kClusterSize - 1) / kClusterSize;
  const long long r0 = (long long)rank * per;
  const long long r1 = (r0 + per < R) ? r0 + per : R;
  const int cb = c0 + g * 8;
  float a0[8], a1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a0[k] = a1[k] = 0.f;
  float sc[8], sh[8], mu[8], is[8];
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = fin.scale[cb + k];
      sh[k] = fin.shift[cb + k];
      mu[k] = fin.mean[cb + k];
      is[k] = fin.invstd[cb + k];
    }
  }
  // dy' = dy masked by the ReLU of the forward output (saved output when a residual was added)
  auto masked = [&](const float (&xf)[8], float (&df)[8], const Bf8& py) {
    if (res_or_mask != nullptr) {
      float yf[8];
      unpack(py, yf);
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = yf[k] > 0.f ? df[k] : 0.f;
    } else if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) df[k] = fmaf(xf[k], sc[k], sh[k]) > 0.f ? df[k] : 0.f;
    }
  };
  constexpr int U = 4;
  for (long long r = r0 + rl; r < r1; r += (long long)U * LANES) {
    Bf8 bx[U], bd[U], by[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const long long rr = r + (long long)j * LANES;
      if (rr < r1) {
        const long long o = rr * C + cb;
        bx[j] = ld8(x + o);
        if (BWD) {
          bd[j] = ld8(dy + o);
          if (res_or_mask != nullptr) by[j] = ld8(res_or_mask + o);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (r + (long long)j * LANES < r1) {
        float xf[8];
        unpack(bx[j], xf);
        if (!BWD) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a0[k] += xf[k];
            a1[k] = fmaf(xf[k], xf[k], a1[k]);
          }
        } else {
          float df[8];
          unpack(bd[j], df);
          masked(xf, df, by[j]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            a0[k] += df[k];
            a1[k] = fmaf(df[k], (xf[k] - mu[k]) * is[k], a1[k]);
          }
        }
      }
    }
  }
  float* mine = red + ((size_t)rl * CGC + g * 8) * 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mine[2 * k] = a0[k];
    mine[2 * k + 1] = a1[k];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < CGC * 2; t += kThreads) {
    float s = 0.f;
    for (int l = 0; l < LANES; ++l) s += red[(size_t)l * CGC * 2 + t];
    part[t] = s;
  }
  cluster_sync_all();                          // every CTA's partial is visible cluster-wide
  if (threadIdx.x < CGC) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (uint32_t k = 0; k < (uint32_t)kClusterSize; ++k) {
      s += (double)ld_dsmem(part + 2 * threadIdx.x, k);
      q += (double)ld_dsmem(part + 2 * threadIdx.x + 1, k);
    }
    const int c = c0 + threadIdx.x;
    if (BWD) {
      float P_, Q_, S_;
      fin.backward(c, C, R, s, q, P_, Q_, S_, rank == 0);
      cs[3 * threadIdx.x] = P_;
      cs[3 * threadIdx.x + 1] = Q_;
      cs[3 * threadIdx.x + 2] = S_;
    } else {
      float sc_, sh_;
      fin.forward(c, R, s, q, sc_, sh_, rank == 0);
      cs[3 * threadIdx.x] = sc_;
      cs[3 * threadIdx.x + 1] = sh_;
    }
  }
  __syncthreads();
  float k0[8], k1[8], k2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    k0[k] = cs[3 * (g * 8 + k)];
    k1[k] = cs[3 * (g * 8 + k) + 1];
    k2[k] = cs[3 * (g * 8 + k) + 2];
  }
  for (long long r = r0 + rl; r < r1; r += (long long)U * LANES) {
    Bf8 bx[U], bd[U], by[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const long long rr = r + (long long)j * LANES;
      if (rr < r1) {
        const long long o = rr * C + cb;
        bx[j] = ld8(x + o);
        if (BWD) bd[j] = ld8(dy + o);
        if (res_or_mask != nullptr) by[j] = ld8(res_or_mask + o);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const long long rr = r + (long long)j * LANES;
      if (rr < r1) {
        const long long o = rr * C + cb;
        float xf[8];
        unpack(bx[j], xf);
        if (!BWD) {
          float rf[8];
          if (res_or_mask != nullptr) {
            unpack(by[j], rf);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) rf[k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float v = fmaf(xf[k], k0[k], k1[k]) + rf[k];
            xf[k] = (relu && v < 0.f) ? 0.f : v;
          }
          st8(out + o, pack(xf));
        } else {
          float df[8];
          unpack(bd[j], df);
          masked(xf, df, by[j]);
          if (dres != nullptr) st8(dres + o, pack(df));
#pragma unroll
          for (int k = 0; k < 8; ++k) df[k] = fmaf(k0[k], df[k], fmaf(k1[k], xf[k], k2[k]));
          st8(out + o, pack(df));
        }
      }
    }
  }
  cluster_sync_all();                          // nobody leaves while a peer may still read its partial
}

bool cluster_mode(long long R, int C) {
  static const bool on = [] {
    const char* e = std::getenv("BYZPY_BN_CLUSTER");
    return !(e && e[0] == '0');
  }();
  static const long long limit = [] {
    const char* e = std::getenv("BYZPY_BN_CLUSTER_MAX_MB");
    return (long long)(e ? atof(e) * (1 << 20) : (4 << 20));
  }();
  return on && (C % 16) == 0 && R * (long long)C * 2 <= limit && R >= kClusterSize;
}

template <int CGC, bool BWD, typename... Args>
cudaError_t launch_cluster(int C, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C / CGC) * kClusterSize);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterSize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, bn_cluster_kernel<CGC, BWD>, args...);
}

int reduce_blocks(long long R, int sm_count) {
  // two resident CTAs per SM saturate HBM with the 8-deep unrolled 16-byte loads; more CTAs only
  // lengthen the finalize
  long long b = (long long)sm_count * 2;
  if (b > R / 64) b = R / 64;
  if (b < 1) b = 1;
  return (int)b;
}

// Small activations take the two-launch (finalize-in-prologue) path with at most 32 partials.
bool fin_mode(long long R, int C) { return R * (long long)C * 2 <= (4ll << 20); }

int fin_blocks(long long R) {
  long long b = R / 64;
  if (b > 32) b = 32;
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
  const auto* res = reinterpret_cast<const __nv_bfloat16*>(a->res);
  const int C = a->C;
  const long long total8 = a->R * (C >> 3);
  const Finalize fin = make_finalize(a);
  int launches = 0;
  bool fused_fin = false;
  int nb = 0;
  if (a->training && cluster_mode(a->R, C)) {
    // one launch: thread-block clusters + distributed shared memory (see bn_cluster_kernel)
    cudaError_t ce = (C <= 512 || (C % 32) != 0)
                         ? launch_cluster<16, false>(C, stream, x, (const __nv_bfloat16*)nullptr, y, a->R, C, a->relu, res,
                                                     (__nv_bfloat16*)nullptr, fin)
                         : launch_cluster<32, false>(C, stream, x, (const __nv_bfloat16*)nullptr, y, a->R, C, a->relu, res,
                                                     (__nv_bfloat16*)nullptr, fin);
    if (ce == cudaSuccess) {
      if (a->launches) *a->launches = 1;
      return (int)cudaGetLastError();
    }
    cudaGetLastError();     // cluster launch unavailable: fall through to the two-launch path
  }
  if (a->training) {
    fused_fin = fin_mode(a->R, C);
    nb = fused_fin ? fin_blocks(a->R) : reduce_blocks(a->R, sm_count);
    const int lanes = kThreads / (C >> 3);
    const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
    reduce_partial_kernel<false><<<nb, kThreads, smem, stream>>>(x, nullptr, a->R, C, nullptr, nullptr, nullptr,
                                                                nullptr, 0, nullptr, a->partial);
    ++launches;
    if (!fused_fin) {
      launch_chain(true, finalize_kernel<false>, (C + 7) / 8, kThreads, 0, stream, (const float*)a->partial, nb, C,
                   a->R, fin);
      ++launches;
    }
  }
  const bool chained = a->training != 0;   // eval mode: nothing of ours precedes the apply kernel
  if (fused_fin) {
    int blocks = stream_blocks(total8, sm_count);
    if (blocks > sm_count) blocks = sm_count;
    launch_chain(chained, apply_kernel<true>, blocks, kThreads, (size_t)C * 2 * sizeof(float), stream, x, y, total8,
                 C, (const float*)a->scale, (const float*)a->shift, a->relu, res, (const float*)a->partial, nb,
                 a->R, fin);
  } else {
    launch_chain(chained, apply_kernel<false>, stream_blocks(total8, sm_count), kThreads, 0, stream, x, y, total8,
                 C, (const float*)a->scale, (const float*)a->shift, a->relu, res, (const float*)nullptr, 0, a->R,
                 fin);
  }
  ++launches;
  if (a->launches) *a->launches = launches;
  return (int)cudaGetLastError();
}

int bz_bn_backward(const BzBnArgs* a, int sm_count, cudaStream_t stream) {
  if (int e = check_shape(a->R, a->C)) return e;
  const auto* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  const auto* dy = reinterpret_cast<const __nv_bfloat16*>(a->dy);
  const auto* ymask = reinterpret_cast<const __nv_bfloat16*>(a->ymask);
  auto* dx = reinterpret_cast<__nv_bfloat16*>(a->dx);
  auto* dres = reinterpret_cast<__nv_bfloat16*>(a->dres);
  const int C = a->C;
  const bool fused_fin = fin_mode(a->R, C);
  const int nb = fused_fin ? fin_blocks(a->R) : reduce_blocks(a->R, sm_count);
  const int lanes = kThreads / (C >> 3);
  const size_t smem = (size_t)lanes * C * 2 * sizeof(float);
  const Finalize fin = make_finalize(a);
  const long long total8 = a->R * (C >> 3);
  int launches = 2;
  if (cluster_mode(a->R, C)) {
    cudaError_t ce = (C <= 512 || (C % 32) != 0)
                         ? launch_cluster<16, true>(C, stream, x, dy, dx, a->R, C, a->relu, ymask, dres, fin)
                         : launch_cluster<32, true>(C, stream, x, dy, dx, a->R, C, a->relu, ymask, dres, fin);
    if (ce == cudaSuccess) {
      if (a->launches) *a->launches = 1;
      return (int)cudaGetLastError();
    }
    cudaGetLastError();
  }
  reduce_partial_kernel<true><<<nb, kThreads, smem, stream>>>(x, dy, a->R, C, a->scale, a->shift, a->mean,
                                                             a->invstd, a->relu, ymask, a->partial);
  if (fused_fin) {
    int blocks = stream_blocks(total8, sm_count);
    if (blocks > sm_count) blocks = sm_count;
    launch_chain(true, bwd_apply_kernel<true>, blocks, kThreads, (size_t)C * 3 * sizeof(float), stream, x, dy, dx,
                 total8, C, (const float*)a->scale, (const float*)a->shift, (const float*)a->coef, a->relu, ymask,
                 dres, (const float*)a->partial, nb, a->R, fin);
  } else {
    launch_chain(true, finalize_kernel<true>, (C + 7) / 8, kThreads, 0, stream, (const float*)a->partial, nb, C,
                 a->R, fin);
    launch_chain(true, bwd_apply_kernel<false>, stream_blocks(total8, sm_count), kThreads, 0, stream, x, dy, dx,
                 total8, C, (const float*)a->scale, (const float*)a->shift, (const float*)a->coef, a->relu, ymask,
                 dres, (const float*)nullptr, 0, a->R, fin);
    launches = 3;
  }
  if (a->launches) *a->launches = launches;
  return (int)cudaGetLastError();
}
