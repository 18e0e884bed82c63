// Single-CTA n-space kernels, part 2: the row maps of the pre-aggregators and the CAF filter.
//
// Like nspace.cu they take the fp64 Gram matrix ON THE DEVICE and leave their result ON THE
// DEVICE, without any host synchronisation, so a pre-aggregated (or CAF) round is a chain of
// launches that a CUDA graph can capture (the host versions in ops/nspace.py stay as oracles).
//
//   clip   W = diag(s),  s_i = min(1, tau / max(||x_i||, 1e-12))      reference pre_aggregators/clipping.py:113-117
//   arc    tau = sorted_norms[n - nb - 1], nb = clamp(floor(2f (n-f) / n), 0, n-1), then like clip
//                                                                       reference pre_aggregators/arc.py:36-51
//   nnm    W_ij = 1/k on the k = n - f rows nearest to x_i (self included, ties to the lower
//          index), squared distances from G                            reference pre_aggregators/nnm.py:82-97
//   caf    covariance-bound agnostic filter in the span of the rows: soft weights c, weighted mean,
//          power iteration on the centred Gram, keep the mean with the smallest top eigenvalue
//                                                                       reference aggregators/norm_wise/caf.py:133-184
#include "api.h"
#include "nspace.h"

namespace {

__device__ __forceinline__ double inf64() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ double norm_of(const double* G, int n, int i) {
  const double g = G[i * n + i];
  if (g != g) return inf64();
  return sqrt(g < 0.0 ? 0.0 : g);
}

// mode 0 clip (param = threshold), 1 arc (iparam = f), 2 nnm (iparam = f).  W is (n, n) fp64
// row-major (and optionally an fp32 copy): the map's matrix, dense because the n-space
// composition G' = W G W^T and the weighted-sum pass both take it that way.
__global__ void __launch_bounds__(1024) preagg_map_kernel(const double* __restrict__ G, int n, int mode,
                                                         double param, int iparam, double* __restrict__ W,
                                                         float* __restrict__ W32) {
  extern __shared__ double sm[];
  double* D = sm;              // n * n (nnm) or n norms
  __shared__ double s_tau;
  const int nn = n * n;
  if (mode == 0 || mode == 1) {
    double* norms = D;
    for (int i = threadIdx.x; i < n; i += blockDim.x) norms[i] = norm_of(G, n, i);
    if (threadIdx.x == 0) s_tau = param;
    __syncthreads();
    if (mode == 1) {
      const int f = iparam;
      long long nb = (long long)floor(2.0 * (double)f / (double)n * (double)(n - f));
      if (nb < 0) nb = 0;
      if (nb > n - 1) nb = n - 1;
      if (threadIdx.x == 0) s_tau = inf64();          // nb == 0: nothing is clipped
      __syncthreads();
      if (nb > 0) {
        const int want = n - (int)nb - 1;              // rank of the threshold norm in ascending order
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const double v = norms[i];
          int rank = 0;
          for (int k = 0; k < n; ++k) rank += (norms[k] < v || (norms[k] == v && k < i)) ? 1 : 0;
          if (rank == want) s_tau = v;
        }
      }
      __syncthreads();
    }
    const double tau = s_tau;
    for (int t = threadIdx.x; t < nn; t += blockDim.x) {
      const int i = t / n, j = t % n;
      double s = 0.0;
      if (i == j) {
        const double nm = norms[i] > 1e-12 ? norms[i] : 1e-12;
        s = tau / nm;                                   // inf / inf -> NaN -> 1 (an all-inf row keeps scale 1)
        if (!(s < 1.0)) s = 1.0;
      }
      W[t] = s;
      if (W32) W32[t] = (float)s;
    }
    return;
  }
  // ---- nnm
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const int i = t / n, j = t % n;
    double d;
    if (i == j) {
      d = -1.0;                                         // self always belongs to its own neighbourhood
    } else {
      d = G[i * n + i] + G[j * n + j] - 2.0 * G[t];
      if (d != d) d = inf64();
      if (d < 0.0) d = 0.0;
    }
    D[t] = d;
  }
  __syncthreads();
  const int k = n - iparam;
  const double wk = 1.0 / (double)k;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const int i = t / n, j = t % n;
    const double* row = D + i * n;
    const double dj = row[j];
    int rank = 0;
    for (int l = 0; l < n; ++l) rank += (row[l] < dj || (row[l] == dj && l < j)) ? 1 : 0;
    const double w = rank < k ? wk : 0.0;
    W[t] = w;
    if (W32) W32[t] = (float)w;
  }
}

// ---------------------------------------------------------------------------------------- CAF
constexpr int kC = 128;   // threads = max rows

__device__ __forceinline__ double csum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < kC / 32; ++w) s += scratch[w];
  return s;
}
__device__ __forceinline__ double cmax(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = scratch[0];
#pragma unroll
  for (int w = 1; w < kC / 32; ++w) s = fmax(s, scratch[w]);
  return s;
}

// G: (n+1, n+1) Gram of the n data rows followed by the start direction r of the power iteration
// (t = X r = G[:n, n], |r|^2 = G[n, n]).  out: n + 1 weights (the last one 0).
// Thread i owns row i; vectors live in shared memory; every matrix-vector product with the centred
// Gram  Gy = Gx - Gp 1^T - 1 Gp^T + (p^T Gp) 1 1^T  is one pass over row i of Gx.
__global__ void __launch_bounds__(kC) caf_kernel(const double* __restrict__ G, int n, int f, int power_iters,
                                                float* __restrict__ out) {
  __shared__ double vec[kC], scratch[kC / 32];
  const int i = threadIdx.x;
  const int ld = n + 1;
  const bool on = i < n;
  const double t_i = on ? G[i * ld + n] : 0.0;
  const double rr = G[n * ld + n];
  const double rnorm = sqrt(rr > 0.0 ? rr : 0.0);
  double c = on ? 1.0 : 0.0;
  double total = (double)n;
  double best = on ? 1.0 / (double)n : 0.0;
  double best_lam = inf64();
  const double stop_at = (double)(n - 2 * f);
  for (int guard = 0; guard < 4 * kC && total > stop_at; ++guard) {
    const double p = c / total;
    // Gp = Gx p
    __syncthreads();
    vec[i] = p;
    __syncthreads();
    double Gp = 0.0;
    if (on)
      for (int j = 0; j < n; ++j) Gp = fma(G[j * ld + i], vec[j], Gp);
    const double pGp = csum(on ? p * Gp : 0.0, scratch);
    const double mu_r = csum(on ? p * t_i : 0.0, scratch);
    double proj = on ? (t_i - mu_r) / (rnorm > 0.0 ? rnorm : 1.0) : 0.0;
    for (int it = 0; it < power_iters; ++it) {
      const double nb = c * proj;
      // Gy nb  (row i)
      __syncthreads();
      vec[i] = on ? nb : 0.0;
      __syncthreads();
      double Gxb = 0.0;
      if (on)
        for (int j = 0; j < n; ++j) Gxb = fma(G[j * ld + i], vec[j], Gxb);
      const double sb = csum(on ? nb : 0.0, scratch);
      const double Gpb = csum(on ? Gp * nb : 0.0, scratch);
      const double Gyb = on ? (Gxb - Gp * sb - Gpb + pGp * sb) : 0.0;
      const double nn2 = csum(on ? nb * Gyb : 0.0, scratch);
      const double nrm = sqrt(nn2 > 0.0 ? nn2 : 0.0);
      if (nrm <= 1e-12) break;                          // uniform across the block
      proj = Gyb / nrm;                                 // Gy (nb / nrm)
    }
    const double csumv = csum(c, scratch);
    const double lam = csum(on ? c * proj * proj : 0.0, scratch) / (csumv > 1e-12 ? csumv : 1e-12);
    if (lam < best_lam) {
      best_lam = lam;
      best = p;
    }
    const double tau = proj * proj;
    const double tau_max = cmax(on ? tau : 0.0, scratch);
    if (tau_max <= 1e-12) break;
    double cn = c * (1.0 - tau / tau_max);
    c = on ? (cn > 0.0 ? cn : 0.0) : 0.0;
    total = csum(c, scratch);
    if (total <= 0.0) break;
  }
  if (on) out[i] = (float)best;
  if (i == n) out[n] = 0.f;
}

}  // namespace

int bz_nspace_preagg(const double* G, int n, int mode, double param, int iparam, double* W, float* W32,
                     cudaStream_t stream) {
  if (n < 1 || n > BZ_MAXN || mode < 0 || mode > 2 || W == nullptr) return (int)cudaErrorInvalidValue;
  if (mode == 2 && !(iparam >= 0 && iparam < n)) return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)(mode == 2 ? n * n : n) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(preagg_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         BZ_MAXN * BZ_MAXN * (int)sizeof(double));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  preagg_map_kernel<<<1, 1024, smem, stream>>>(G, n, mode, param, iparam, W, W32);
  return (int)cudaGetLastError();
}

int bz_nspace_caf(const double* G, int n, int f, int power_iters, float* out, cudaStream_t stream) {
  if (n < 1 || n >= kC || 2 * f >= n || f < 0 || power_iters < 0) return (int)cudaErrorInvalidValue;
  caf_kernel<<<1, kC, 0, stream>>>(G, n, f, power_iters, out);
  return (int)cudaGetLastError();
}
