// tcgen05 / TMEM / TMA Gram kernel (gram_umma.cu).
#pragma once
#include "api.h"

struct BzGramUmmaArgs {
  RowTable rows;
  ScaleTable scales;
  int n;
  int n_pad;               // filled by the launcher: 16 / 32 / 64 / 128
  long long off, len;      // the kernel consumes floor(len / tile_cols(n)) * tile_cols(n) columns starting at off
  float* partials;         // num_partials x 2 x n x n floats (num_partials >= partials(n, grid))
  int num_partials;
  const double* tail64;    // optional (n, n) fp64 Gram of the remaining columns, added in the reduce
  float* G;
  double* G64;             // optional
};

int bz_gram_umma_npad(int n);
int bz_gram_umma_tile_cols(int n);          // columns consumed per tile: 32 * (128 / n_pad)
int bz_gram_umma_grid(int n, long long len, int sm_count);
int bz_gram_umma_partials(int n, int grid);  // partial slots written by `grid` CTAs
int bz_gram_umma(const BzGramUmmaArgs* args, int sm_count, cudaStream_t stream);
