// Argument block of the fused cross-GPU parameter-server kernels.
#pragma once
#include "api.h"

#define BZ_MAXW 8            // ranks on one NVSwitch box
#define BZ_PAD_READY 0       // pad[BZ_PAD_READY + src] : src's gradient rows are ready (epoch)
#define BZ_PAD_DONE 16       // pad[BZ_PAD_DONE  + src] : src delivered its shard (epoch)
#define BZ_PAD_GRAM 32       // pad[BZ_PAD_GRAM  + src] : src delivered its partial Gram (epoch)
#define BZ_PAD_WORDS 64      // uint32 words per signal pad

struct BzFusedPsArgs {
  RowTable rows;       // n gradient rows: local or peer-mapped addresses
  ScaleTable scales;
  int n;
  VirtRows virt;
  int f;
  int mode;            // BzCwMode
  long long d;         // padded arena length (multiple of 4)
  long long shard_off, shard_len;  // this rank's coordinate shard (multiples of 4)
  int rank, world;
  float* agg[BZ_MAXW];      // aggregated-gradient buffer of every rank (peer-mapped)
  uint32_t* pad[BZ_MAXW];   // signal pad of every rank (peer-mapped)
  uint32_t epoch;
  const uint32_t* epoch_ptr;  // optional device-resident epoch (CUDA-graph replay); overrides `epoch`
  unsigned int* counter;    // local CTA arrival counter (zero-initialised)
  int* status;              // local error word (0 == ok)
  UpdTable upd;             // local replicas to update in phase 2
  int grid_limit;           // 0 = one full co-resident wave; >0 caps the CTA count
};

int bz_fused_ps_cw(const BzFusedPsArgs* args, int sm_count, cudaStream_t stream);
int bz_bump_u32(uint32_t* p, cudaStream_t stream);
