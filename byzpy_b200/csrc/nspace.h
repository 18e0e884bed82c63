// Launchers of the single-CTA n-space solvers (nspace.cu).
#pragma once
#include <cuda_runtime.h>

int bz_nspace_krum(const double* G, int n, int f, int q, float* w, cudaStream_t stream);
int bz_nspace_weiszfeld(const double* G, int nt, int n_real, const double* a0, double tol,
                        int max_iter, double eps, float* out, int* iters, cudaStream_t stream);
int bz_nspace_cclip(const double* G, int nt, int n_real, const double* a0, double c_tau, int M,
                    double eps, float* out, cudaStream_t stream);
