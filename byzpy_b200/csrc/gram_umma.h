// tcgen05 / TMEM / TMA Gram kernel (gram_umma.cu).
#pragma once
#include "api.h"

struct BzGramUmmaArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  long long off, len;      // the kernel consumes floor(len / 32) * 32 columns starting at off
  float* partials;         // num_partials x 2 x n x n floats
  int num_partials;
  const double* tail64;    // optional (n, n) fp64 Gram of the remaining columns, added in the reduce
  float* G;
  double* G64;             // optional
};

int bz_gram_umma_grid(long long len, int sm_count);
int bz_gram_umma(const BzGramUmmaArgs* args, int sm_count, cudaStream_t stream);
