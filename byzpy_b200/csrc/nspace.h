// Launchers of the single-CTA n-space solvers (nspace.cu).
#pragma once
#include <cuda_runtime.h>

int bz_nspace_krum(const double* G, int n, int f, int q, float* w, cudaStream_t stream);
int bz_nspace_weiszfeld(const double* G, int nt, int n_real, const double* a0, double tol,
                        int max_iter, double eps, float* out, int* iters, cudaStream_t stream);
int bz_nspace_cclip(const double* G, int nt, int n_real, const double* a0, double c_tau, int M,
                    double eps, float* out, cudaStream_t stream);

// Exhaustive (n - f)-subset search for MDA (mode 0) / SMEA (mode 1) on the device, n <= 24.
// G is fp64 with leading dimension ldg; w gets 1/m on the winning rows (nt entries, rest 0).
// scratch_score / scratch_rank hold bz_nspace_subset_blocks(n, m, sm_count) entries.
unsigned long long bz_binomial(int n, int k);
int bz_nspace_subset_blocks(int n, int m, int sm_count);
int bz_nspace_subset(const double* G, int ldg, int n, int m, int nt, int mode, double* scratch_score,
                     unsigned long long* scratch_rank, float* w, int sm_count, cudaStream_t stream);

// nspace_maps.cu -------------------------------------------------------------------------------
// Row map of a pre-aggregator as an (n, n) fp64 matrix on the device: mode 0 clip (param =
// threshold), 1 ARC (iparam = f), 2 nearest-neighbour mixing (iparam = f).  W32 (optional): fp32 copy.
int bz_nspace_preagg(const double* G, int n, int mode, double param, int iparam, double* W, float* W32,
                     cudaStream_t stream);
// CAF filter on the (n+1, n+1) Gram of [rows; start direction]; out: n+1 weights.
int bz_nspace_caf(const double* G, int n, int f, int power_iters, float* out, cudaStream_t stream);
