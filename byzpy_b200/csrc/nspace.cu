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
