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
