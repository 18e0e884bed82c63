// Gram matrix  G = (S X)(S X)^T  on CUDA cores, exact fp32 products with a
// deterministic two-stage split-K reduction (fp64 final accumulation).
//
// This is the reference-precision path and the fallback for the tcgen05
// kernel in gram_umma.cu.  "Pass 1" of every Gram-family operator (Krum,
// Multi-Krum, MoNNA, CGE, NNM, ARC, Clipping, MDA, SMEA, and the Gram-space
// Weiszfeld / centered-clipping / CAF solves): one read of n*d*4
// bytes, (n, n) result.
// Parity: reference krum.py:31-44 (pairwise squared distances via Gram).
#include <cstdlib>

#include "api.h"
#include "cw_core.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---------------------------------------------------------------- small n --
// n <= 16: every thread streams V consecutive coordinates of all rows and keeps
// the upper triangle of the outer product in registers.
// AUX: one more row = the coordinate-wise lower median of the (scaled) rows, computed in registers
// with the selection network of cw_core.cuh, stored to a.aux_median and included in the products.
// STAGED (experiment, BYZPY_GRAM_SMALL_IMPL=2): every thread streams its next tiles into a thread-private
// slot of a 3-deep shared-memory ring with cp.async (the staging helpers of cw_core.cuh).  Measured slower
// than the direct register form (profiles/gram_small.md), kept for A/B runs.
constexpr int kGramStages = 3;

template <int NS, int V, bool AUX, bool STAGED>
__global__ void __launch_bounds__(kThreads) gram_small_kernel(const __grid_constant__ BzGramArgs a) {
  extern __shared__ __align__(16) float gram_stage_mem[];
  constexpr int NR = NS + (AUX ? 1 : 0);
  constexpr int T = NR * (NR + 1) / 2;
  float acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.f;
  const int n = a.n;
  const int ne = n + (AUX ? 1 : 0);
  const int apad = NS / 2 - 1 - (n - 1) / 2;
  const long long nvec = a.len / V;
  const long long stride = (long long)gridDim.x * kThreads;
  auto accumulate = [&](const float (&col)[NR]) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int k = i; k < NR; ++k) {
        acc[t] = fmaf(col[i], col[k], acc[t]);
        ++t;
      }
  };
  auto median_of = [&](const float (&col)[NR]) -> float {
    float v[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = (i < n) ? canon(col[i] * a.scales.s[i]) : 0.f;
    return bzcw::cw_pick<NS, BZ_CW_MEDIAN>(v, n, 0, apad);
  };
  const size_t stage_elems = (size_t)n * kThreads * V;
  const long long u0 = (long long)blockIdx.x * kThreads + threadIdx.x;
  if constexpr (STAGED) {
#pragma unroll
    for (int st = 0; st < kGramStages - 1; ++st) {
      const long long u = u0 + st * stride;
      if (u < nvec) bzcw::cw_stage_issue<NS, V>(gram_stage_mem + st * stage_elems, kThreads, a.rows, n, a.off + u * V);
      bzcw::cp_async_commit();
    }
  }
  int slot = 0;
  for (long long u = u0; u < nvec; u += stride) {
    const long long base = a.off + u * V;
    float x[NS][V];
    if constexpr (STAGED) {
      const long long un = u + (kGramStages - 1) * stride;
      int sn = slot + kGramStages - 1;
      if (sn >= kGramStages) sn -= kGramStages;
      if (un < nvec) bzcw::cw_stage_issue<NS, V>(gram_stage_mem + sn * stage_elems, kThreads, a.rows, n, a.off + un * V);
      bzcw::cp_async_commit();
      bzcw::cp_async_wait<kGramStages - 1>();
      const float* stage = gram_stage_mem + slot * stage_elems;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        if (i < n) {
          const float* p = stage + ((size_t)i * kThreads + threadIdx.x) * V;
          if constexpr (V == 4) {
            const float4 t = *reinterpret_cast<const float4*>(p);
            x[i][0] = t.x; x[i][1] = t.y; x[i][2] = t.z; x[i][3] = t.w;
          } else if constexpr (V == 2) {
            const float2 t = *reinterpret_cast<const float2*>(p);
            x[i][0] = t.x; x[i][1] = t.y;
          } else {
            x[i][0] = p[0];
          }
        } else {
#pragma unroll
          for (int c = 0; c < V; ++c) x[i][c] = 0.f;
        }
      }
      if (++slot == kGramStages) slot = 0;
    } else {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      if (i < n) {
        if constexpr (V == 4) {
          const float4 t = ldg_stream4(a.rows.p[i] + base);
          x[i][0] = t.x; x[i][1] = t.y; x[i][2] = t.z; x[i][3] = t.w;
        } else if constexpr (V == 2) {
          const float2 t = ldg_stream2(a.rows.p[i] + base);
          x[i][0] = t.x; x[i][1] = t.y;
        } else {
          x[i][0] = ldg_stream1(a.rows.p[i] + base);
        }
      } else {
#pragma unroll
        for (int c = 0; c < V; ++c) x[i][c] = 0.f;
      }
    }
    }
    float med[V];
#pragma unroll
    for (int c = 0; c < V; ++c) {
      float col[NR];
#pragma unroll
      for (int i = 0; i < NS; ++i) col[i] = x[i][c];
      if constexpr (AUX) {
        med[c] = median_of(col);
        col[NS] = med[c];
      }
      accumulate(col);
    }
    if constexpr (AUX) {
      if constexpr (V == 4) {
        stg_stream4(a.aux_median + base, make_float4(med[0], med[1], med[2], med[3]));
      } else {
#pragma unroll
        for (int c = 0; c < V; ++c) a.aux_median[base + c] = med[c];
      }
    }
  }
  if constexpr (STAGED) bzcw::cp_async_wait<0>();
  // scalar tail (len % V) handled by block 0
  if (V > 1 && blockIdx.x == 0) {
    const long long tail0 = a.off + nvec * V;
    const long long tail = a.len - nvec * V;
    if ((long long)threadIdx.x < tail) {
      float col[NR];
#pragma unroll
      for (int i = 0; i < NS; ++i) col[i] = (i < n) ? a.rows.p[i][tail0 + threadIdx.x] : 0.f;
      if constexpr (AUX) {
        col[NS] = median_of(col);
        a.aux_median[tail0 + threadIdx.x] = col[NS];
      }
      accumulate(col);
    }
  }
  __shared__ float red[kWarps][T];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float s = warp_sum(acc[t]);
    if (lane == 0) red[warp][t] = s;
  }
  __syncthreads();
  float* part = a.partials + (size_t)blockIdx.x * ne * ne;
  for (int t = threadIdx.x; t < T; t += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][t];
    // decode t -> (i, k), i <= k  (register rows; the aux row NS is logical row n)
    int i = 0, rem = t;
    while (rem >= NR - i) {
      rem -= NR - i;
      ++i;
    }
    int k = i + rem;
    if (AUX && i == NS) i = n;
    else if (i >= n) continue;
    if (AUX && k == NS) k = n;
    else if (k >= n) continue;
    part[i * ne + k] = s;
    part[k * ne + i] = s;
  }
}

// ---------------------------------------------------------------- tiled n --
// 16 < n <= 128: 32-column tiles staged (transposed) in shared memory with a
// register prefetch double buffer; 16x16 threads each own an RB x RB block.
template <int RB>
__global__ void __launch_bounds__(kThreads) gram_tiled_kernel(const __grid_constant__ BzGramArgs a) {
  constexpr int NPAD = 16 * RB, TJ = 32, LD = NPAD + 1;
  constexpr int RPW = NPAD / kWarps;  // rows per warp
  __shared__ float Xs[2][TJ][LD];
  const int n = a.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long nchunks = (a.len + TJ - 1) / TJ;
  float acc[RB][RB];
#pragma unroll
  for (int p = 0; p < RB; ++p)
#pragma unroll
    for (int q = 0; q < RB; ++q) acc[p][q] = 0.f;

  float pref[RPW];
  auto gload = [&](long long chunk) {
    const long long col = chunk * TJ + lane;
    const bool ok = col < a.len;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int i = warp + kWarps * r;
      pref[r] = (ok && i < n) ? ldg_stream1(a.rows.p[i] + a.off + col) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) Xs[buf][lane][warp + kWarps * r] = pref[r];
  };

  long long chunk = blockIdx.x;
  int buf = 0;
  if (chunk < nchunks) {
    gload(chunk);
    sstore(0);
  }
  __syncthreads();
  for (; chunk < nchunks; chunk += gridDim.x) {
    const long long next = chunk + gridDim.x;
    if (next < nchunks) gload(next);
#pragma unroll 8
    for (int j = 0; j < TJ; ++j) {
      float av[RB], bv[RB];
#pragma unroll
      for (int p = 0; p < RB; ++p) av[p] = Xs[buf][j][ty + 16 * p];
#pragma unroll
      for (int q = 0; q < RB; ++q) bv[q] = Xs[buf][j][tx + 16 * q];
#pragma unroll
      for (int p = 0; p < RB; ++p)
#pragma unroll
        for (int q = 0; q < RB; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    if (next < nchunks) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  float* part = a.partials + (size_t)blockIdx.x * n * n;
#pragma unroll
  for (int p = 0; p < RB; ++p)
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const int i = ty + 16 * p, k = tx + 16 * q;
      if (i < n && k < n) part[i * n + k] = acc[p][q];
    }
}

__global__ void gram_reduce_kernel(const float* __restrict__ partials, int num_partials, int n,
                                   ScaleTable scales, float* __restrict__ G,
                                   double* __restrict__ G64) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  double s = 0.0;
  for (int c = 0; c < num_partials; ++c) s += (double)partials[(size_t)c * n * n + t];
  const int i = t / n, k = t % n;
  s *= (double)scales.s[i] * (double)scales.s[k];
  G[t] = (float)s;
  if (G64) G64[t] = s;
}

int grid_for(int n, long long len, int sm_count) {
  long long blocks;
  if (n <= 16) {
    const int V = (n <= 8) ? 4 : 2;
    const long long nvec = len / V;
    blocks = (nvec + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count * 4;
    if (blocks > cap) blocks = cap;
  } else {
    blocks = (len + 31) / 32;
    const long long cap = (long long)sm_count * 2;
    if (blocks > cap) blocks = cap;
  }
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int bz_gram_partials_needed(int n, int sm_count) { return sm_count * 4 * n * n; }

int bz_gram(const BzGramArgs* args, int sm_count, cudaStream_t stream) {
  const BzGramArgs& a0 = *args;
  if (a0.n < 1 || a0.n > BZ_MAXN || a0.len < 0) return (int)cudaErrorInvalidValue;
  BzGramArgs a = a0;
  const int n = a.n;
  const int grid = grid_for(n, a.len, sm_count);
  if (grid > a.num_partials) return (int)cudaErrorInvalidValue;
  if (a.aux_median != nullptr && n >= BZ_MAXN) return (int)cudaErrorInvalidValue;
  bool al16 = (a.off % 4) == 0, al8 = (a.off % 2) == 0;
  for (int i = 0; i < n; ++i) {
    al16 = al16 && ((uintptr_t)a.rows.p[i] % 16) == 0;
    al8 = al8 && ((uintptr_t)a.rows.p[i] % 8) == 0;
  }
  const bool aux = a.aux_median != nullptr;
  if (aux) {
    if (n > 16) return (int)cudaErrorInvalidValue;
    al16 = al16 && ((uintptr_t)a.aux_median % 16) == 0;
    al8 = al8 && ((uintptr_t)a.aux_median % 8) == 0;
    a.scales.s[n] = 1.f;           // the median row is built from already-scaled values
  }
  // the cp.async-staged form pays off once every thread has several tiles to stream (impl: 0 auto,
  // 1 direct, 2 staged -- BYZPY_GRAM_SMALL_IMPL, for A/B measurements)
  static const int forced = [] {
    const char* e = getenv("BYZPY_GRAM_SMALL_IMPL");
    return e ? atoi(e) : 0;
  }();
#define BZ_GRAM_LAUNCH(NS_, V_, AUX_, ST_)                                                          \
  do {                                                                                              \
    const size_t smem = ST_ ? (size_t)kGramStages * n * kThreads * V_ * sizeof(float) : 0;          \
    if (ST_) {                                                                                      \
      static bool conf = false;                                                                     \
      if (!conf) {                                                                                  \
        cudaError_t ce = cudaFuncSetAttribute(gram_small_kernel<NS_, V_, AUX_, ST_>,                \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); \
        if (ce != cudaSuccess) return (int)ce;                                                      \
        conf = true;                                                                                \
      }                                                                                             \
    }                                                                                               \
    gram_small_kernel<NS_, V_, AUX_, ST_><<<grid, kThreads, smem, stream>>>(a);                     \
  } while (0)
#define BZ_GRAM_SMALL(NS_, V_)                                                                      \
  do {                                                                                              \
    /* measured (profiles/gram_small.md): the staged form is SLOWER here (n = 8: 0.21 vs 0.17 ms for 537 MB, */ \
    /* n = 16: 0.49 vs 0.38 ms) -- the extra shared-memory round trip costs more than the deeper queue wins */ \
    /* -- so it only runs when forced */                                                                  \
    const bool st_ = NS_ >= 8 && forced == 2;                                                       \
    if (aux) {                                                                                      \
      if (st_) BZ_GRAM_LAUNCH(NS_, V_, true, true);                                                 \
      else BZ_GRAM_LAUNCH(NS_, V_, true, false);                                                    \
    } else {                                                                                        \
      if (st_) BZ_GRAM_LAUNCH(NS_, V_, false, true);                                                \
      else BZ_GRAM_LAUNCH(NS_, V_, false, false);                                                   \
    }                                                                                               \
  } while (0)
  if (n <= 2) {
    if (al16) BZ_GRAM_SMALL(2, 4);
    else BZ_GRAM_SMALL(2, 1);
  } else if (n <= 4) {
    if (al16) BZ_GRAM_SMALL(4, 4);
    else BZ_GRAM_SMALL(4, 1);
  } else if (n <= 8) {
    if (al16) BZ_GRAM_SMALL(8, 4);
    else BZ_GRAM_SMALL(8, 1);
  } else if (n <= 16) {
    if (al8) BZ_GRAM_SMALL(16, 2);
    else BZ_GRAM_SMALL(16, 1);
#undef BZ_GRAM_SMALL
#undef BZ_GRAM_LAUNCH
  } else if (n <= 32) {
    gram_tiled_kernel<2><<<grid, kThreads, 0, stream>>>(a);
  } else if (n <= 64) {
    gram_tiled_kernel<4><<<grid, kThreads, 0, stream>>>(a);
  } else {
    gram_tiled_kernel<8><<<grid, kThreads, 0, stream>>>(a);
  }
  int e = (int)cudaGetLastError();
  if (e) return e;
  const int rt = 128;
  const int ne = n + (aux ? 1 : 0);
  gram_reduce_kernel<<<(ne * ne + rt - 1) / rt, rt, 0, stream>>>(a.partials, grid, ne, a.scales, a.G,
                                                                 a.G64);
  return (int)cudaGetLastError();
}
