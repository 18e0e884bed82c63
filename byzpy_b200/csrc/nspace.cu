// Single-CTA n-space solvers (n <= 128 rows + a few auxiliary rows).
//
// Input: the fp64 Gram matrix ON THE DEVICE (output of gram.cu / gram_umma.cu).
// Output: a weight / coefficient vector ON THE DEVICE, consumed by wsum.cu.
// No host synchronisation anywhere, so a Gram-family aggregation is three
// back-to-back launches; data-dependent loops (Weiszfeld's stopping rule) are
// resolved inside the kernel.
//
// Semantics mirror byzpy_b200/ops/nspace.py (the host oracle):
//   krum       scores = sum of the n-f-1 smallest off-diagonal squared distances,
//              q best scores get weight 1/q            (reference krum.py:177-194)
//   weiszfeld  a <- w / sum(w), w_i = 1/max(||x_i - z||, eps), stop when
//              ||z_new - z|| <= tol                     (reference geometric_median.py:87-102)
//   cclip      a <- (1 - s/n) a + alpha/n, alpha_i = min(1, tau/max(||x_i - v||, eps))
//                                                       (reference center_clipping.py:146-154)
// Distances to the iterate z = sum_j a_j x_j come from G alone:
//   ||x_i - z||^2 = G_ii - 2 (G a)_i + a^T G a.
#include "api.h"
#include "nspace.h"

namespace {

constexpr int kT = 160;  // >= BZ_MAXN + auxiliary rows (129), multiple of 32

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  // full-block sum, result broadcast to every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kT / 32; ++w) s += scratch[w];
  return s;
}

__device__ __forceinline__ double sqdist_entry(const double* G, int nt, int i, int j) {
  if (i == j) return 0.0;
  double d = G[i * nt + i] + G[j * nt + j] - 2.0 * G[i * nt + j];
  if (d != d) return __longlong_as_double(0x7ff0000000000000LL);
  return d < 0.0 ? 0.0 : d;
}

// 1024 threads; D (n x n, fp64) and the scores live in dynamic shared memory.
//   phase 1: D_ij from the Gram matrix;
//   phase 2: one (i, j) pair per thread-iteration: rank of D_ij inside row i by counting
//            (ties -> lower index first), entries of rank 1 .. n-f-1 are added to score_i;
//   phase 3: rank of score_i among the scores; the q best get weight 1/q.
__global__ void __launch_bounds__(1024) krum_kernel(const double* __restrict__ G, int n, int f, int q,
                                                   float* __restrict__ w) {
  extern __shared__ double sm[];
  double* D = sm;                 // n * n
  double* score = sm + n * n;     // n
  const int nn = n * n;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) D[t] = sqdist_entry(G, n, t / n, t % n);
  for (int t = threadIdx.x; t < n; t += blockDim.x) score[t] = 0.0;
  __syncthreads();
  const int keep = n - f;         // sorted positions 1 .. n-f-1 are summed (position 0 = self)
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const int i = t / n, j = t % n;
    const double* row = D + i * n;
    const double dj = row[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const double dk = row[k];
      rank += (dk < dj || (dk == dj && k < j)) ? 1 : 0;
    }
    if (rank >= 1 && rank < keep) atomicAdd(&score[i], dj);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = score[i];
    int rank = 0;
    for (int k = 0; k < n; ++k) rank += (score[k] < s || (score[k] == s && k < i)) ? 1 : 0;
    w[i] = rank < q ? (float)(1.0 / (double)q) : 0.f;
  }
}

// (G v)_i for thread i; G symmetric so column access G[j*nt + i] is coalesced.
__device__ __forceinline__ double matvec_row(const double* G, int nt, int i, const double* v) {
  double s = 0.0;
  for (int j = 0; j < nt; ++j) s = fma(G[j * nt + i], v[j], s);
  return s;
}

__device__ __forceinline__ double dist_to_iterate(const double* G, int nt, int i, double Ga_i,
                                                  double q) {
  double d2 = G[i * nt + i] - 2.0 * Ga_i + q;
  if (d2 != d2) return __longlong_as_double(0x7ff0000000000000LL);
  return sqrt(d2 < 0.0 ? 0.0 : d2);
}

__global__ void __launch_bounds__(kT) weiszfeld_kernel(const double* __restrict__ G, int nt,
                                                      int n_real, const double* __restrict__ a0,
                                                      double tol, int max_iter, double eps,
                                                      float* __restrict__ out,
                                                      int* __restrict__ iters_out) {
  __shared__ double a[kT], dl[kT], scratch[kT / 32];
  const int i = threadIdx.x;
  a[i] = (i < nt) ? a0[i] : 0.0;
  __syncthreads();
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    const double Ga = (i < nt) ? matvec_row(G, nt, i, a) : 0.0;
    const double q = block_sum((i < nt) ? a[i] * Ga : 0.0, scratch);
    double wi = 0.0;
    if (i < n_real) {
      const double dist = dist_to_iterate(G, nt, i, Ga, q);
      wi = 1.0 / (dist > eps ? dist : eps);
    }
    const double W = block_sum(wi, scratch);
    const double an = wi / W;
    __syncthreads();
    dl[i] = (i < nt) ? an - a[i] : 0.0;
    __syncthreads();
    const double Gd = (i < nt) ? matvec_row(G, nt, i, dl) : 0.0;
    const double step2 = block_sum((i < nt) ? dl[i] * Gd : 0.0, scratch);
    __syncthreads();
    a[i] = an;
    __syncthreads();
    const double step = sqrt(step2 < 0.0 ? 0.0 : step2);
    if (!(step > tol)) break;  // also exits on NaN, like the host oracle
  }
  if (i < nt) out[i] = (float)a[i];
  if (i == 0 && iters_out) *iters_out = it > max_iter ? max_iter : it;
}

__global__ void __launch_bounds__(kT) cclip_kernel(const double* __restrict__ G, int nt, int n_real,
                                                  const double* __restrict__ a0, double c_tau, int M,
                                                  double eps, float* __restrict__ out) {
  __shared__ double a[kT], scratch[kT / 32];
  const int i = threadIdx.x;
  a[i] = (i < nt) ? a0[i] : 0.0;
  __syncthreads();
  for (int it = 0; it < M; ++it) {
    const double Ga = (i < nt) ? matvec_row(G, nt, i, a) : 0.0;
    const double q = block_sum((i < nt) ? a[i] * Ga : 0.0, scratch);
    double alpha = 0.0;
    if (i < n_real) {
      const double dist = dist_to_iterate(G, nt, i, Ga, q);
      const double r = c_tau / (dist > eps ? dist : eps);
      alpha = r < 1.0 ? r : 1.0;
    }
    const double s = block_sum(alpha, scratch);
    __syncthreads();
    a[i] = (1.0 - s / (double)n_real) * a[i] + alpha / (double)n_real;
    __syncthreads();
  }
  if (i < nt) out[i] = (float)a[i];
}

// ---------------------------------------------------------------------------------------------
// Exhaustive subset search on the device (MDA / SMEA, reference minimum_diameter_average.py:358-386
// and smea.py:63-88): all C(n, m) subsets of size m = n - f are enumerated in itertools.combinations
// (lexicographic) order, one subset per thread-iteration, unranked with the combinatorial number
// system from a binomial table in shared memory.
//   MODE 0 (MDA)   score = max pairwise squared distance inside the subset
//   MODE 1 (SMEA)  score = top eigenvalue of the centred m x m Gram block / m  (cyclic Jacobi)
// The best (score, rank) pair -- ties go to the smaller rank, i.e. the lexicographically first
// subset, like the host oracle -- is reduced per CTA, then across CTAs by a second tiny kernel that
// also decodes the winner into the weight vector.  No host synchronisation: the selection is
// CUDA-graph capturable.
constexpr int kSubN = 24;        // rows handled by the exhaustive search (C(24,12) = 2.7 M subsets)
constexpr int kSubThreads = 128;

__device__ __forceinline__ void unrank_subset(unsigned long long rank, int n, int m,
                                              const unsigned long long* binom /* [kSubN+1][kSubN+1] */,
                                              int* idx) {
  // lexicographic unranking: choose the smallest next element whose block contains `rank`
  int x = 0;
  for (int i = 0; i < m; ++i) {
    for (;; ++x) {
      const unsigned long long cnt = binom[(n - x - 1) * (kSubN + 1) + (m - i - 1)];
      if (rank < cnt) break;
      rank -= cnt;
    }
    idx[i] = x++;
  }
}

__device__ double jacobi_top_eigenvalue(double* A, int m) {
  // cyclic Jacobi on the symmetric m x m matrix A (thread-private, row-major); returns max eigenvalue
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) off += A[p * m + q] * A[p * m + q];
    if (off < 1e-26) break;
    for (int p = 0; p < m; ++p) {
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[p * m + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * m + q] - A[p * m + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < m; ++k) {
          const double akp = A[k * m + p], akq = A[k * m + q];
          A[k * m + p] = c * akp - sn * akq;
          A[k * m + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {
          const double apk = A[p * m + k], aqk = A[q * m + k];
          A[p * m + k] = c * apk - sn * aqk;
          A[q * m + k] = sn * apk + c * aqk;
        }
      }
    }
  }
  double top = A[0];
  for (int i = 1; i < m; ++i) top = fmax(top, A[i * m + i]);
  return top;
}

template <int MODE>
__global__ void __launch_bounds__(kSubThreads) subset_search_kernel(const double* __restrict__ G, int ldg, int n,
                                                                   int m, unsigned long long total,
                                                                   double* __restrict__ best_score,
                                                                   unsigned long long* __restrict__ best_rank) {
  __shared__ unsigned long long binom[(kSubN + 1) * (kSubN + 1)];
  __shared__ double M[kSubN * kSubN];          // MDA: squared distances; SMEA: the Gram block itself
  __shared__ double rs[kSubThreads];
  __shared__ unsigned long long rr[kSubThreads];
  for (int t = threadIdx.x; t < (kSubN + 1) * (kSubN + 1); t += kSubThreads) {
    const int a = t / (kSubN + 1), b = t % (kSubN + 1);
    // C(a, b) by the multiplicative formula (exact in 64 bits for a <= 24)
    unsigned long long c = (b > a) ? 0ull : 1ull;
    if (b <= a)
      for (int i = 1; i <= b; ++i) c = c * (unsigned long long)(a - b + i) / (unsigned long long)i;
    binom[t] = c;
  }
  for (int t = threadIdx.x; t < n * n; t += kSubThreads) {
    const int i = t / n, j = t % n;
    M[t] = (MODE == 0) ? sqdist_entry(G, ldg, i, j) : G[i * ldg + j];
  }
  __syncthreads();
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double my_score = inf;
  unsigned long long my_rank = ~0ull;
  double A[(MODE == 1) ? kSubN * kSubN : 1];
  const unsigned long long stride = (unsigned long long)gridDim.x * kSubThreads;
  for (unsigned long long r = (unsigned long long)blockIdx.x * kSubThreads + threadIdx.x; r < total; r += stride) {
    int idx[kSubN];
    unrank_subset(r, n, m, binom, idx);
    double score;
    if (MODE == 0) {
      score = 0.0;
      for (int a = 0; a < m; ++a)
        for (int b = a + 1; b < m; ++b) score = fmax(score, M[idx[a] * n + idx[b]]);
      if (score != score) score = inf;
    } else {
      // centred block: H S H with H = I - 11^T/m  ->  s_ab - rowmean_a - rowmean_b + mean
      double rowm[kSubN], tot = 0.0;
      for (int a = 0; a < m; ++a) {
        double acc = 0.0;
        for (int b = 0; b < m; ++b) acc += M[idx[a] * n + idx[b]];
        rowm[a] = acc / m;
        tot += acc;
      }
      tot /= (double)m * m;
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) A[a * m + b] = M[idx[a] * n + idx[b]] - rowm[a] - rowm[b] + tot;
      for (int a = 0; a < m; ++a)
        for (int b = a + 1; b < m; ++b) A[a * m + b] = A[b * m + a] = 0.5 * (A[a * m + b] + A[b * m + a]);
      const double top = jacobi_top_eigenvalue(A, m);
      score = (top != top) ? inf : fmax(top, 0.0) / m;
    }
    if (score < my_score || (score == my_score && r < my_rank)) {
      my_score = score;
      my_rank = r;
    }
  }
  rs[threadIdx.x] = my_score;
  rr[threadIdx.x] = my_rank;
  __syncthreads();
  for (int o = kSubThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double s2 = rs[threadIdx.x + o];
      const unsigned long long r2 = rr[threadIdx.x + o];
      if (s2 < rs[threadIdx.x] || (s2 == rs[threadIdx.x] && r2 < rr[threadIdx.x])) {
        rs[threadIdx.x] = s2;
        rr[threadIdx.x] = r2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    best_score[blockIdx.x] = rs[0];
    best_rank[blockIdx.x] = rr[0];
  }
}

__global__ void subset_pick_kernel(const double* __restrict__ best_score,
                                   const unsigned long long* __restrict__ best_rank, int nblocks, int n, int m,
                                   int nt, float* __restrict__ w) {
  __shared__ unsigned long long binom[(kSubN + 1) * (kSubN + 1)];
  for (int t = threadIdx.x; t < (kSubN + 1) * (kSubN + 1); t += blockDim.x) {
    const int a = t / (kSubN + 1), b = t % (kSubN + 1);
    unsigned long long c = (b > a) ? 0ull : 1ull;
    if (b <= a)
      for (int i = 1; i <= b; ++i) c = c * (unsigned long long)(a - b + i) / (unsigned long long)i;
    binom[t] = c;
  }
  for (int t = threadIdx.x; t < nt; t += blockDim.x) w[t] = 0.f;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double s = best_score[0];
  unsigned long long r = best_rank[0];
  for (int b = 1; b < nblocks; ++b) {
    if (best_score[b] < s || (best_score[b] == s && best_rank[b] < r)) {
      s = best_score[b];
      r = best_rank[b];
    }
  }
  int idx[kSubN];
  if (r == ~0ull) {   // every score was non-finite: the first m rows, like the host oracle
    for (int i = 0; i < m; ++i) idx[i] = i;
  } else {
    unrank_subset(r, n, m, binom, idx);
  }
  for (int i = 0; i < m; ++i) w[idx[i]] = 1.f / (float)m;
}

}  // namespace

int bz_nspace_krum(const double* G, int n, int f, int q, float* w, cudaStream_t stream) {
  if (n < 1 || n > BZ_MAXN || f < 0 || q < 1) return (int)cudaErrorInvalidValue;
  const int smem = (n * n + n) * (int)sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(krum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (BZ_MAXN * BZ_MAXN + BZ_MAXN) * (int)sizeof(double));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  krum_kernel<<<1, 1024, smem, stream>>>(G, n, f, q, w);
  return (int)cudaGetLastError();
}

int bz_nspace_weiszfeld(const double* G, int nt, int n_real, const double* a0, double tol,
                        int max_iter, double eps, float* out, int* iters, cudaStream_t stream) {
  if (nt < 1 || nt > kT || n_real < 1 || n_real > nt) return (int)cudaErrorInvalidValue;
  weiszfeld_kernel<<<1, kT, 0, stream>>>(G, nt, n_real, a0, tol, max_iter, eps, out, iters);
  return (int)cudaGetLastError();
}

int bz_nspace_cclip(const double* G, int nt, int n_real, const double* a0, double c_tau, int M,
                    double eps, float* out, cudaStream_t stream) {
  if (nt < 1 || nt > kT || n_real < 1 || n_real > nt) return (int)cudaErrorInvalidValue;
  cclip_kernel<<<1, kT, 0, stream>>>(G, nt, n_real, a0, c_tau, M, eps, out);
  return (int)cudaGetLastError();
}

// Host-side binomial (saturating) so callers can decide whether the exhaustive search is feasible.
unsigned long long bz_binomial(int n, int k) {
  if (k < 0 || k > n) return 0;
  if (k > n - k) k = n - k;
  unsigned long long c = 1;
  for (int i = 1; i <= k; ++i) {
    const unsigned long long num = (unsigned long long)(n - k + i);
    if (c > (~0ull) / num) return ~0ull;
    c = c * num / (unsigned long long)i;
  }
  return c;
}

int bz_nspace_subset_blocks(int n, int m, int sm_count) {
  const unsigned long long total = bz_binomial(n, m);
  unsigned long long b = (total + kSubThreads - 1) / kSubThreads;
  const unsigned long long cap = (unsigned long long)sm_count * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int bz_nspace_subset(const double* G, int ldg, int n, int m, int nt, int mode, double* scratch_score,
                     unsigned long long* scratch_rank, float* w, int sm_count, cudaStream_t stream) {
  if (n < 1 || n > kSubN || m < 1 || m > n || nt < n || ldg < n || (mode != 0 && mode != 1))
    return (int)cudaErrorInvalidValue;
  const unsigned long long total = bz_binomial(n, m);
  if (total == ~0ull) return (int)cudaErrorInvalidValue;
  const int blocks = bz_nspace_subset_blocks(n, m, sm_count);
  if (mode == 0)
    subset_search_kernel<0><<<blocks, kSubThreads, 0, stream>>>(G, ldg, n, m, total, scratch_score, scratch_rank);
  else
    subset_search_kernel<1><<<blocks, kSubThreads, 0, stream>>>(G, ldg, n, m, total, scratch_score, scratch_rank);
  subset_pick_kernel<<<1, 128, 0, stream>>>(scratch_score, scratch_rank, blocks, n, m, nt, w);
  return (int)cudaGetLastError();
}
