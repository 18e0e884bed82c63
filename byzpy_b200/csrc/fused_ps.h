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
  long long shard_off, shard_len;  // this rank's coordinate shard of the launch's range (multiples of 4)
  // Coordinate range [rng_off, rng_off + rng_len) this launch aggregates and updates: one gradient
  // BUCKET.  A round is a sequence of bucket launches in reverse-layer order, each enqueued as soon as
  // the local backward pass has produced the bucket, so aggregation overlaps the rest of backward.
  long long rng_off, rng_len;
  int rank, world;
  uint32_t live_mask;       // bit r set: rank r takes part (0 = all `world` ranks); silent ranks are
                            // neither waited for nor written to
  float* agg[BZ_MAXW];      // aggregated-gradient buffer of every rank (peer-mapped)
  float* agg_mc;            // NVLS multicast alias of the agg buffers (one multimem.st reaches every
                            // rank's HBM through the switch); nullptr -> world peer stores
  uint32_t* pad[BZ_MAXW];   // signal pad of every rank (peer-mapped)
  uint32_t epoch;
  const uint32_t* epoch_ptr;  // optional device-resident epoch (CUDA-graph replay); overrides `epoch`
  // Flag words carry the monotone sequence number  epoch * seq_mul + seq_add  (bucket b of a round
  // with nb buckets: seq_mul = nb, seq_add = b), so one word per (kind, rank) serves every bucket.
  uint32_t seq_mul, seq_add;
  unsigned long long spin_ns;  // wall-clock budget of every flag wait (0 = default 20 s)
  // Optional device-side timeline of this launch (nullptr = off): 8 %globaltimer stamps [ns] --
  // 0 kernel start, 1 ready wait over, 2 block 0 finished phase 1, 3 last CTA finished phase 1,
  // 4 delivery wait over, 5 block 0 finished the optimizer step (utils/tracing.py renders them).
  unsigned long long* trace;
  unsigned int* counter;    // local CTA arrival counter (zero-initialised)
  int* status;              // local error word (0 == ok)
  UpdTable upd;             // local replicas to update in phase 2
  int grid_limit;           // 0 = one full co-resident wave; >0 caps the CTA count
  const float* W;           // wsum mode: n weights on the device (output of the n-space solve)
};

int bz_fused_ps_cw(const BzFusedPsArgs* args, int sm_count, cudaStream_t stream);
int bz_bump_u32(uint32_t* p, cudaStream_t stream);
// *dst = %globaltimer [ns] when the stream reaches this point (round timelines, utils/tracing.py)
int bz_stamp(unsigned long long* dst, cudaStream_t stream);

// Gram-family round pieces -------------------------------------------------------------
// Fused  Y = W (S X)  on this rank's shard + broadcast + SGD (same protocol as bz_fused_ps_cw).
int bz_fused_ps_wsum(const BzFusedPsArgs* args, int sm_count, cudaStream_t stream);

// Device-side flag barrier: publish pad[slot + rank] = epoch on every rank, wait for all.
struct BzFlagBarrierArgs {
  uint32_t* pad[BZ_MAXW];
  int rank, world;
  int slot;                  // BZ_PAD_READY / BZ_PAD_GRAM / ...
  const uint32_t* epoch_ptr;
  int* status;
  uint32_t live_mask;        // see BzFusedPsArgs
  uint32_t seq_mul, seq_add; // flag value = epoch * seq_mul + seq_add (seq_mul 0 -> plain epoch)
  unsigned long long spin_ns;
};
int bz_flag_barrier(const BzFlagBarrierArgs* args, cudaStream_t stream);

// All-reduce of the (n, n) fp64 partial Gram through peer stores: write my partial into
// slot[rank] of every rank, flag, wait, sum the `world` slots locally.
struct BzGramExchangeArgs {
  const double* local;       // (n, n) partial of this rank
  double* slots[BZ_MAXW];    // per rank: world x n x n doubles (peer-mapped)
  uint32_t* pad[BZ_MAXW];
  int rank, world, n;
  const uint32_t* epoch_ptr;
  int* status;
  uint32_t live_mask;
  unsigned long long spin_ns;
  double* out64;             // (n, n) total
  float* out32;              // (n, n) total (optional)
  // NVLS form (slots_mc != nullptr): slot 0 of every rank holds its partial, slot 1 receives the total.
  // Rank r reduces its slice of the matrix INSIDE THE SWITCH with multimem.ld_reduce.add.f64 over the
  // team's slot-0 copies and multimem.st's the sums into everybody's slot 1 -- one load and one store
  // per element instead of world peer stores + world local loads, and every rank receives bit-identical
  // totals (each element is summed once, by one rank), which the n-space selections rely on.
  double* slots_mc;
};
int bz_gram_exchange(const BzGramExchangeArgs* args, cudaStream_t stream);
