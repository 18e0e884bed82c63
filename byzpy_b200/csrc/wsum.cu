// Weighted row-sum  Y = W (S X)  for m <= 8 output rows.
//
// This is "pass 2" of every Gram-family aggregator: after the n-space solve
// produced a tiny weight matrix W (m x n) ON THE DEVICE, one streaming pass
// over the n row buffers (local or peer HBM) produces the m output vectors.
// Row scales (clipping factors, sign flips) are folded into W in shared
// memory, so pre-aggregated / attacked vectors are never materialised.
//   Krum/Multi-Krum/MoNNA/CGE/GeometricMedian/CenteredClipping/CAF -> m = 1
// Parity: reference krum.py:192-194, monna.py:81-82, geometric_median.py:96-98,
//         center_clipping.py:146-154, bucketing.py:115-119 (semantics only).
#include "api.h"

namespace {

constexpr int kThreads = 256;

template <int V>
__device__ __forceinline__ void sgd_apply_w(const UpdTable& upd, long long idx,
                                            const float (&g)[V]) {
  for (int r = 0; r < upd.count; ++r) {
    float* pp = upd.param[r] + idx;
    float* mp = upd.mom[r] ? upd.mom[r] + idx : nullptr;
#pragma unroll
    for (int c = 0; c < V; ++c) {
      const float p = pp[c];
      float gg = g[c] + upd.wd * p;
      if (mp) {
        const float mo = upd.mu * mp[c] + gg;
        mp[c] = mo;
        gg = mo;
      }
      pp[c] = p - upd.lr * gg;
    }
  }
}

template <int M, int V>
__global__ void __launch_bounds__(kThreads) wsum_kernel(const __grid_constant__ BzWsumArgs a) {
  __shared__ float Ws[M][BZ_MAXN];
  const int n = a.n;
  for (int t = threadIdx.x; t < M * BZ_MAXN; t += kThreads) {
    const int r = t / BZ_MAXN, i = t % BZ_MAXN;
    Ws[r][i] = (r < a.m && i < n) ? a.W[r * n + i] * a.scales.s[i] : 0.f;
  }
  __syncthreads();
  const long long nvec = a.len / V;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
    const long long base = a.off + u * V;
    float acc[M][V];
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
      for (int c = 0; c < V; ++c) acc[r][c] = 0.f;
    // 8 independent 16-byte loads in flight per thread
    for (int i0 = 0; i0 < n; i0 += 8) {
      float x[8][V];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k;
        if (i < n) {
          if constexpr (V == 4) {
            const float4 t = ldg_stream4(a.rows.p[i] + base);
            x[k][0] = t.x; x[k][1] = t.y; x[k][2] = t.z; x[k][3] = t.w;
          } else {
            x[k][0] = ldg_stream1(a.rows.p[i] + base);
          }
        } else {
#pragma unroll
          for (int c = 0; c < V; ++c) x[k][c] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k;
        if (i < n) {
#pragma unroll
          for (int r = 0; r < M; ++r) {
            const float w = Ws[r][i];
            // w == 0 rows are skipped so that +-inf/NaN in an excluded
            // (e.g. Byzantine) row cannot poison the output with 0*inf.
            if (w != 0.f) {
#pragma unroll
              for (int c = 0; c < V; ++c) acc[r][c] = fmaf(w, x[k][c], acc[r][c]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < M; ++r) {
      if (r < a.m && a.out[r] != nullptr) {
        if constexpr (V == 4) {
          stg_stream4(a.out[r] + base, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
        } else {
          a.out[r][base] = acc[r][0];
        }
      }
    }
    if (a.upd.count > 0) sgd_apply_w<V>(a.upd, base, acc[0]);
  }
}

template <int M, int V>
int launch_one(const BzWsumArgs& a, int sm_count, cudaStream_t stream) {
  const long long nvec = a.len / V;
  if (nvec <= 0) return 0;
  long long blocks = (nvec + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  wsum_kernel<M, V><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int M>
int launch_m(const BzWsumArgs& a, bool vec_ok, int sm_count, cudaStream_t stream) {
  if (vec_ok) {
    const long long main_len = a.len - (a.len % 4);
    BzWsumArgs b = a;
    b.len = main_len;
    int e = launch_one<M, 4>(b, sm_count, stream);
    if (e) return e;
    if (main_len < a.len) {
      b.off = a.off + main_len;
      b.len = a.len - main_len;
      e = launch_one<M, 1>(b, sm_count, stream);
    }
    return e;
  }
  return launch_one<M, 1>(a, sm_count, stream);
}

}  // namespace

int bz_wsum(const BzWsumArgs* args, int sm_count, cudaStream_t stream) {
  const BzWsumArgs& a = *args;
  if (a.n < 1 || a.n > BZ_MAXN || a.m < 1 || a.m > 8) return (int)cudaErrorInvalidValue;
  bool vec_ok = (a.off % 4) == 0;
  for (int i = 0; i < a.n && vec_ok; ++i) vec_ok = ((uintptr_t)a.rows.p[i] % 16) == 0;
  for (int r = 0; r < a.m && vec_ok; ++r)
    vec_ok = a.out[r] == nullptr || ((uintptr_t)a.out[r] % 16) == 0;
  if (a.m == 1) return launch_m<1>(a, vec_ok, sm_count, stream);
  if (a.m == 2) return launch_m<2>(a, vec_ok, sm_count, stream);
  if (a.m <= 4) return launch_m<4>(a, vec_ok, sm_count, stream);
  return launch_m<8>(a, vec_ok, sm_count, stream);
}

// ------------------------------------------------------------------------------------------
// Y = W (S X) for 8 < m <= 128 output rows in ONE pass over the inputs (NNM emits n mixed vectors,
// Bucketing n / s bucket means; reference pre_aggregators/nnm.py:92-93 does this as `mask @ flat`).
// The m <= 8 kernel above re-reads the n x d matrix once per 8 output rows; here a CTA stages a
// [n x 128] coordinate tile in shared memory with cp.async (two stages) next to the whole weight
// matrix (transposed, scales folded in) and computes the [m x 128] output tile as a register-tiled
// fp32 GEMM: 16 x 16 threads, each owning RM rows x 8 coordinates (two 4-wide groups, bank-conflict
// free), 3 LDS.128 per 8 RM FMAs.  The
// arithmetic (m n FMAs per coordinate) is the bound: at m = n = 64 it is 2x the single-pass byte time
// instead of the 4.5x of eight passes.  Zero weights are skipped by predication, like above, so a
// +-inf / NaN row excluded from an output cannot poison it with 0 * inf.
namespace {

constexpr int kTileC = 128;          // coordinates per tile
constexpr int kMultiStages = 2;

template <int RM>
__global__ void __launch_bounds__(kThreads) wsum_multi_kernel(const __grid_constant__ BzWsumMultiArgs a) {
  extern __shared__ __align__(16) float wm_smem[];
  const int n = a.n, m = a.m;
  constexpr int MP = 16 * RM;                       // padded output rows
  float* Wt = wm_smem;                              // [n][MP]  (transposed: Wt[k][r] = W[r][k] * scale[k])
  float* Xs = wm_smem + (size_t)n * MP;             // kMultiStages x [n][kTileC]
  for (int t = threadIdx.x; t < n * MP; t += kThreads) {
    const int k = t / MP, r = t % MP;
    Wt[t] = (r < m) ? a.W[(size_t)r * n + k] * a.scales.s[k] : 0.f;
  }
  const long long ntiles = a.len / kTileC;
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  const size_t stage_elems = (size_t)n * kTileC;
  auto issue = [&](long long tile, int stage) {
    const long long col = a.off + tile * kTileC;
    float* dst = Xs + stage * stage_elems;
    // n rows x 32 chunks of 16 bytes
    for (int q = threadIdx.x; q < n * (kTileC / 4); q += kThreads) {
      const int k = q >> 5, ch = q & 31;
      const unsigned sd = (unsigned)__cvta_generic_to_shared(dst + (size_t)k * kTileC + ch * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd), "l"(a.rows.p[k] + col + ch * 4) : "memory");
    }
  };
  long long tile = blockIdx.x;
  if (tile < ntiles) issue(tile, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int stage = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < ntiles) issue(next, stage ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                  // the tile (and, first time, Wt) is visible to everyone
    const float* X = Xs + stage * stage_elems;
    float acc[RM][8];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
      // coordinates tc*4 .. +3 and 64 + tc*4 .. +3: consecutive threads read consecutive 16-byte chunks
      // (a 32-byte stride per thread would put threads 0/4/8/12 on the same banks: 4-way conflict)
      const float4 x0 = *reinterpret_cast<const float4*>(X + (size_t)k * kTileC + tc * 4);
      const float4 x1 = *reinterpret_cast<const float4*>(X + (size_t)k * kTileC + 64 + tc * 4);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float wv[RM];
      if constexpr (RM >= 4) {
#pragma unroll
        for (int i = 0; i < RM; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(Wt + (size_t)k * MP + tr * RM + i);
          wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < RM; ++i) wv[i] = Wt[(size_t)k * MP + tr * RM + i];
      }
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        if (wv[i] != 0.f) {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(wv[i], xv[c], acc[i][c]);
        }
      }
    }
    const long long base = a.off + tile * kTileC + tc * 4;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int r = tr * RM + i;
      if (r < m) {
        stg_stream4(a.out.p[r] + base, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        stg_stream4(a.out.p[r] + base + 64, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
      }
    }
    __syncthreads();                                  // everyone is done with this stage before it is refilled
    stage ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int RM>
int launch_multi(const BzWsumMultiArgs& a, int sm_count, cudaStream_t stream) {
  const size_t smem = ((size_t)a.n * 16 * RM + (size_t)kMultiStages * a.n * kTileC) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wsum_multi_kernel<RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsum_multi_kernel<RM>, kThreads, smem) != cudaSuccess || occ < 1)
    occ = 1;
  long long blocks = a.len / kTileC;
  const long long cap = (long long)sm_count * occ;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  wsum_multi_kernel<RM><<<(unsigned)blocks, kThreads, smem, stream>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

int bz_wsum_multi_tile() { return kTileC; }

int bz_wsum_multi(const BzWsumMultiArgs* args, int sm_count, cudaStream_t stream) {
  const BzWsumMultiArgs& a = *args;
  if (a.n < 1 || a.n > BZ_MAXN || a.m < 1 || a.m > BZ_MAXN || a.W == nullptr) return (int)cudaErrorInvalidValue;
  if ((a.off % 4) != 0 || (a.len % kTileC) != 0) return (int)cudaErrorInvalidValue;
  for (int i = 0; i < a.n; ++i)
    if (((uintptr_t)a.rows.p[i] % 16) != 0) return (int)cudaErrorInvalidValue;
  for (int r = 0; r < a.m; ++r)
    if (a.out.p[r] == nullptr || ((uintptr_t)a.out.p[r] % 16) != 0) return (int)cudaErrorInvalidValue;
  if (a.m <= 16) return launch_multi<1>(a, sm_count, stream);
  if (a.m <= 32) return launch_multi<2>(a, sm_count, stream);
  if (a.m <= 64) return launch_multi<4>(a, sm_count, stream);
  return launch_multi<8>(a, sm_count, stream);
}
