// tcgen05 / TMEM / TMA Gram kernel (gram_umma.cu).
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is reached through dlopen, see vmm.cpp)

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

// ---- TMA-fed variant -------------------------------------------------------------------------
// When the n rows form a few row-major matrix SEGMENTS (a stacked (n, d) tensor, a flat arena, or
// one (L, d_pad) gradient matrix per peer GPU in the fused round), the raw fp32 tiles are brought
// into shared memory by the tensor memory accelerator: one cp.async.bulk.tensor.2d per segment
// per tile (box = segment rows x tile columns), completion counted on an mbarrier, issued by one
// elected thread.  Pointer tables of unrelated rows keep the per-thread cp.async path.
#define BZ_GRAM_MAXSEG 12
struct alignas(64) BzGramTmaMaps {
  CUtensorMap maps[BZ_GRAM_MAXSEG];   // 2-D fp32 maps: dim0 = columns (row length), dim1 = segment rows
  int nseg;
  int seg_row0[BZ_GRAM_MAXSEG];       // first Gram row of the segment
  int seg_rows[BZ_GRAM_MAXSEG];
};
int bz_gram_umma_tma(const BzGramUmmaArgs* args, const BzGramTmaMaps* maps, int sm_count, cudaStream_t stream);
// Encode one 2-D tiled fp32 tensor map (driver API through dlopen; returns a CUresult-style code,
// -1 when the driver is not available).  box = box_rows x box_cols elements, no swizzle.
int bz_encode_map_2d(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols,
                     unsigned long long row_stride_bytes, unsigned box_rows, unsigned box_cols);
