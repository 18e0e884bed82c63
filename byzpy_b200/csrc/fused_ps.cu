// Fused cross-GPU parameter-server round for the coordinate-wise family.
//
// ONE kernel launch per gradient BUCKET per rank replaces  all_gather -> aggregate ->
// broadcast -> optimizer.step()  of the reference round (reference
// engine/parameter_server/ps.py:103-144) with no NCCL call on the path.  A round is a short
// sequence of bucket launches in reverse-layer order (parallel/device_ps.py): the launch of
// bucket b is enqueued as soon as the local backward pass has produced that coordinate range, so
// gather / selection / broadcast / SGD of the big late layers run while backward is still working
// on the early ones.  For the launch's coordinate range [rng_off, rng_off + rng_len):
//
//   phase 0  publish "my gradient rows of this bucket are ready" to every peer's signal pad
//            (st.release.sys over NVLink);
//   phase 1  this rank owns the shard [shard_off, shard_off+shard_len) of the range: for each
//            tile it LOADS the n gradient rows directly from the owning GPUs' HBM (16-byte P2P
//            loads through NVSwitch), runs the register selection network (median / trimmed
//            mean / mean-of-medians, attack rows folded in), and delivers the aggregated tile to
//            every rank's `agg` buffer -- ONE multimem.st through the NVLS multicast alias when
//            the symmetric heap has one, `world` P2P stores otherwise;
//   phase 2  after every peer signalled "my shard is delivered", SGD(+momentum) on all local
//            model replicas over the bucket range from the local `agg` buffer.
//
// Synchronisation is device-side only: monotonically increasing sequence numbers
// (epoch * buckets + bucket) in per-rank signal pads, written with release semantics at system
// scope and polled with acquire loads.  Safety argument (no host barrier between rounds):
//   * a peer can only overwrite my agg range for the next round after it saw my ready flag for
//     that round and bucket, which I publish after my previous round (its SGD reads) finished;
//   * I only finish a bucket launch after all peers' done flags, which they publish after they
//     finished reading my gradient rows of the bucket, so my next backward may overwrite them.
// Every spin loop has a wall-clock budget; on expiry the kernel records
//   status |= code | (mask of the ranks that did not arrive) << 8     (code bits: 1 ready wait,
//   2 delivery wait, 4 flag barrier, 8 Gram exchange)
// and exits instead of hanging the GPU; the host side turns that into an exception or drops the
// silent ranks (DeviceRound.recover).
#include "cw_core.cuh"
#include "fused_ps.h"

namespace {
using namespace bzcw;

constexpr unsigned long long kSpinBudgetNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void stamp(unsigned long long* trace, int slot) {
  if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) trace[slot] = globaltimer_ns();
}
__device__ __forceinline__ bool is_live(uint32_t live_mask, int r) {
  return live_mask == 0u || ((live_mask >> r) & 1u) != 0u;
}

// Block-wide wait until flags[r] >= seq for every live rank r.  Returns false on timeout or when
// another wait of this rank already failed (uniform across the block).
__device__ bool wait_all(const uint32_t* flags, int world, uint32_t live_mask, uint32_t seq, int* status,
                         int code, unsigned long long budget) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < world && is_live(live_mask, threadIdx.x)) {
    const unsigned long long t0 = globaltimer_ns();
    if (budget == 0ull) budget = kSpinBudgetNs;
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - seq) < 0) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > budget) {
        s_ok = 0;
        atomicOr(status, code | (1 << (8 + threadIdx.x)));
        break;
      }
      if (*((volatile int*)status) != 0) {
        s_ok = 0;
        break;
      }
    }
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  __syncthreads();
  return ok;
}

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Publish pad[p][slot + rank] = seq on every live rank p.  Called by ALL threads of a CTA (warp 0
// does the work): lane p issues ONE system-scope fence followed by a relaxed store to peer p -- a
// release pattern per lane, so the `world` flag stores leave in parallel behind a single MEMBAR.SYS
// of the warp.  (One thread looping over st.release.sys paid a fence per peer: measured 2.7 us per
// peer on an idle fabric and 15 us per peer while the peers' phase-1 traffic was in flight, i.e.
// >100 us to reach the last rank -- profiles/round_timeline.md.)
__device__ __forceinline__ void publish(uint32_t* const* pad, int slot, int rank, int world, uint32_t live_mask,
                                        uint32_t seq) {
  if ((int)threadIdx.x < world && is_live(live_mask, threadIdx.x)) {
    __threadfence_system();
    st_relaxed_sys(pad[threadIdx.x] + slot + rank, seq);
  }
}

// Deliver V aggregated coordinates to every live rank's agg buffer.
template <int V>
__device__ __forceinline__ void deliver(const BzFusedPsArgs& a, long long base, const float (&res)[V]) {
  if (a.agg_mc != nullptr) {
    float* dst = a.agg_mc + base;
    if constexpr (V == 4) {
      stg_multimem4(dst, make_float4(res[0], res[1], res[2], res[3]));
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) stg_multimem1(dst + c, res[c]);
    }
    return;
  }
  for (int p = 0; p < a.world; ++p) {
    if (!is_live(a.live_mask, p)) continue;
    float* dst = a.agg[p] + base;
    if constexpr (V == 4) {
      stg_stream4(dst, make_float4(res[0], res[1], res[2], res[3]));
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) stg_stream1(dst + c, res[c]);
    }
  }
}

// Phase 1 epilogue + phase 2: the last CTA of this rank announces delivery, then every CTA applies
// the optimizer step over the launch's coordinate range once all ranks have delivered.
__device__ __forceinline__ void finish_round(const BzFusedPsArgs& a, uint32_t seq, uint32_t* my_pad) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();           // this CTA's deliveries are performed system-wide before it is counted
    const unsigned int prev = atomicAdd(a.counter, 1u);
    s_last = (prev == gridDim.x - 1);
    if (s_last) *a.counter = 0u;
  }
  __syncthreads();
  if (s_last) {
    // last CTA of this rank: my shard has been delivered everywhere
    publish(a.pad, BZ_PAD_DONE, a.rank, a.world, a.live_mask, seq);
    if (a.trace != nullptr && threadIdx.x == 0) a.trace[3] = globaltimer_ns();
  }
  stamp(a.trace, 2);
  if (a.upd.count <= 0 && a.world == 1) return;
  if (!wait_all(my_pad + BZ_PAD_DONE, a.world, a.live_mask, seq, a.status, 2, a.spin_ns)) return;
  stamp(a.trace, 4);
  if (a.upd.count > 0) {
    const float* agg = a.agg[a.rank];
    const long long nvec4 = a.rng_len / 4;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec4; u += stride) {
      const long long j = a.rng_off + u * 4;
      const float4 g4 = ldg_cg4(agg + j);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      sgd_apply<4>(a.upd, j, g);
    }
    const long long j = a.rng_off + nvec4 * 4 + (long long)blockIdx.x * kThreads + threadIdx.x;
    if (j < a.rng_off + a.rng_len) {
      const float g[1] = {ldg_cg1(agg + j)};
      sgd_apply<1>(a.upd, j, g);
    }
  }
  stamp(a.trace, 5);
}

__device__ __forceinline__ uint32_t flag_seq(const BzFusedPsArgs& a) {
  const uint32_t epoch = a.epoch_ptr ? *a.epoch_ptr : a.epoch;
  return a.seq_mul ? epoch * a.seq_mul + a.seq_add : epoch;
}

template <int NP, int V, int MODE>
__global__ void __launch_bounds__(kThreads) fused_ps_cw_kernel(const __grid_constant__ BzFusedPsArgs a) {
  uint32_t* my_pad = a.pad[a.rank];
  const uint32_t seq = flag_seq(a);
  // ---- phase 0: publish readiness of my gradient rows (this bucket) -----------
  stamp(a.trace, 0);
  if (blockIdx.x == 0) publish(a.pad, BZ_PAD_READY, a.rank, a.world, a.live_mask, seq);
  // ---- phase 1: gather + select + broadcast my shard ---------------------------
  if (!wait_all(my_pad + BZ_PAD_READY, a.world, a.live_mask, seq, a.status, 1, a.spin_ns)) return;
  stamp(a.trace, 1);
  {
    const long long nvec = a.shard_len / V;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
      const long long base = a.shard_off + u * V;
      float res[V];
      // peers may have been writing these rows after this kernel became resident: no .nc loads
      cw_unit<NP, V, MODE, /*NC=*/false>(a.rows, a.scales, a.n, a.virt, a.f, base, res);
      deliver<V>(a, base, res);
    }
  }
  finish_round(a, seq, my_pad);
}

// ---------------------------------------------------------------- Gram-family pass 2 --
__global__ void __launch_bounds__(kThreads) fused_ps_wsum_kernel(const __grid_constant__ BzFusedPsArgs a) {
  __shared__ float ws[BZ_MAXN];
  uint32_t* my_pad = a.pad[a.rank];
  const uint32_t seq = flag_seq(a);
  const int n = a.n;
  for (int i = threadIdx.x; i < BZ_MAXN; i += kThreads) ws[i] = (i < n) ? a.W[i] * a.scales.s[i] : 0.f;
  stamp(a.trace, 0);
  if (blockIdx.x == 0) publish(a.pad, BZ_PAD_READY, a.rank, a.world, a.live_mask, seq);
  if (!wait_all(my_pad + BZ_PAD_READY, a.world, a.live_mask, seq, a.status, 1, a.spin_ns)) return;
  stamp(a.trace, 1);
  {
    const long long nvec = a.shard_len / 4;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
      const long long base = a.shard_off + u * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i0 = 0; i0 < n; i0 += 8) {
        float4 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = i0 + k;
          x[k] = (i < n && ws[i] != 0.f) ? ldg_weak4(a.rows.p[i] + base) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = i0 + k;
          if (i < n) {
            const float w = ws[i];
            if (w != 0.f) {
              acc[0] = fmaf(w, x[k].x, acc[0]);
              acc[1] = fmaf(w, x[k].y, acc[1]);
              acc[2] = fmaf(w, x[k].z, acc[2]);
              acc[3] = fmaf(w, x[k].w, acc[3]);
            }
          }
        }
      }
      deliver<4>(a, base, acc);
    }
  }
  finish_round(a, seq, my_pad);
}

__global__ void flag_barrier_kernel(const __grid_constant__ BzFlagBarrierArgs a) {
  const uint32_t epoch = *a.epoch_ptr;
  const uint32_t seq = a.seq_mul ? epoch * a.seq_mul + a.seq_add : epoch;
  publish(a.pad, a.slot, a.rank, a.world, a.live_mask, seq);
  wait_all(a.pad[a.rank] + a.slot, a.world, a.live_mask, seq, a.status, 4, a.spin_ns);
}

__global__ void gram_exchange_kernel(const __grid_constant__ BzGramExchangeArgs a) {
  const uint32_t epoch = *a.epoch_ptr;
  const int nn = a.n * a.n;
  for (int p = 0; p < a.world; ++p) {
    if (!is_live(a.live_mask, p)) continue;
    double* dst = a.slots[p] + (size_t)a.rank * nn;
    for (int t = threadIdx.x; t < nn; t += blockDim.x) dst[t] = a.local[t];
  }
  __syncthreads();
  publish(a.pad, BZ_PAD_GRAM, a.rank, a.world, a.live_mask, epoch);
  if (!wait_all(a.pad[a.rank] + BZ_PAD_GRAM, a.world, a.live_mask, epoch, a.status, 8, a.spin_ns)) return;
  const double* mine = a.slots[a.rank];
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) {
      if (!is_live(a.live_mask, r)) continue;
      double v;
      asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(mine + (size_t)r * nn + t));
      s += v;
    }
    a.out64[t] = s;
    if (a.out32) a.out32[t] = (float)s;
  }
}

// (n, n) fp64 all-reduce through the NVLS multicast alias (see BzGramExchangeArgs::slots_mc).
__global__ void __launch_bounds__(1024) gram_exchange_nvls_kernel(const __grid_constant__ BzGramExchangeArgs a) {
  const uint32_t epoch = *a.epoch_ptr;
  const int nn = a.n * a.n;
  double* mine = a.slots[a.rank];            // slot 0: my partial, slot 1 (at + nn): the total
  for (int t = threadIdx.x; t < nn; t += blockDim.x) mine[t] = a.local[t];
  __syncthreads();
  publish(a.pad, BZ_PAD_GRAM, a.rank, a.world, a.live_mask, 2u * epoch);
  if (!wait_all(a.pad[a.rank] + BZ_PAD_GRAM, a.world, a.live_mask, 2u * epoch, a.status, 8, a.spin_ns)) return;
  // my slice of the matrix: elements t with (t / 32) % live == my index among the live ranks
  int live = 0, idx = 0;
  for (int r = 0; r < a.world; ++r) {
    if (!is_live(a.live_mask, r)) continue;
    if (r == a.rank) idx = live;
    ++live;
  }
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    if (((t >> 5) % live) != idx) continue;
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(a.slots_mc + t) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(a.slots_mc + nn + t), "d"(v) : "memory");
  }
  __syncthreads();
  publish(a.pad, BZ_PAD_GRAM, a.rank, a.world, a.live_mask, 2u * epoch + 1u);
  if (!wait_all(a.pad[a.rank] + BZ_PAD_GRAM, a.world, a.live_mask, 2u * epoch + 1u, a.status, 8, a.spin_ns)) return;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(mine + nn + t));
    a.out64[t] = v;
    if (a.out32) a.out32[t] = (float)v;
  }
}

template <int NP, int MODE>
int launch_np(const BzFusedPsArgs& a, int grid, cudaStream_t stream) {
  constexpr int V = (NP <= 16) ? 4 : (NP == 32 ? 2 : 1);
  fused_ps_cw_kernel<NP, V, MODE><<<grid, kThreads, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int NP, int MODE>
int max_grid_np(int sm_count) {
  constexpr int V = (NP <= 16) ? 4 : (NP == 32 ? 2 : 1);
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, fused_ps_cw_kernel<NP, V, MODE>, kThreads, 0);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  return per_sm * sm_count;
}

template <int MODE>
int dispatch_mode(const BzFusedPsArgs& a, int sm_count, cudaStream_t stream, bool query) {
  const int nt = a.n + a.virt.count;
#define BZ_CASE(NP)                                                   \
  if (nt <= NP) {                                                     \
    int g = max_grid_np<NP, MODE>(sm_count);                          \
    if (a.grid_limit > 0 && a.grid_limit < g) g = a.grid_limit;       \
    if (query) return g;                                              \
    return launch_np<NP, MODE>(a, g, stream);                         \
  }
  BZ_CASE(2) BZ_CASE(4) BZ_CASE(8) BZ_CASE(16) BZ_CASE(32) BZ_CASE(64) BZ_CASE(128)
#undef BZ_CASE
  return query ? 0 : (int)cudaErrorInvalidValue;
}

__global__ void bump_u32_kernel(uint32_t* p) { *p = *p + 1u; }
__global__ void stamp_kernel(unsigned long long* dst) { *dst = globaltimer_ns(); }

// normalise the optional fields of the argument block (callers from before buckets existed)
BzFusedPsArgs normalised(const BzFusedPsArgs& in) {
  BzFusedPsArgs a = in;
  if (a.rng_len <= 0) {
    a.rng_off = 0;
    a.rng_len = a.d;
  }
  return a;
}

bool ranges_ok(const BzFusedPsArgs& a) {
  if ((a.shard_off % 4) != 0 || (a.shard_len % 4) != 0 || (a.rng_off % 4) != 0) return false;
  if (a.rng_off < 0 || a.rng_off + a.rng_len > a.d) return false;
  if (a.shard_len > 0 && (a.shard_off < a.rng_off || a.shard_off + a.shard_len > a.rng_off + a.rng_len))
    return false;
  if (a.agg_mc && ((uintptr_t)a.agg_mc % 16) != 0) return false;
  return true;
}

}  // namespace

int bz_bump_u32(uint32_t* p, cudaStream_t stream) {
  bump_u32_kernel<<<1, 1, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

int bz_stamp(unsigned long long* dst, cudaStream_t stream) {
  stamp_kernel<<<1, 1, 0, stream>>>(dst);
  return (int)cudaGetLastError();
}

int bz_fused_ps_wsum(const BzFusedPsArgs* args, int sm_count, cudaStream_t stream) {
  const BzFusedPsArgs a = normalised(*args);
  if (a.n < 1 || a.n > BZ_MAXN || a.W == nullptr || a.world < 1 || a.world > BZ_MAXW)
    return (int)cudaErrorInvalidValue;
  if (!ranges_ok(a) || (a.rng_len % 4) != 0) return (int)cudaErrorInvalidValue;
  for (int i = 0; i < a.n; ++i)
    if (((uintptr_t)a.rows.p[i] % 16) != 0) return (int)cudaErrorInvalidValue;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_ps_wsum_kernel, kThreads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int g = per_sm * sm_count;
  if (a.grid_limit > 0 && a.grid_limit < g) g = a.grid_limit;
  fused_ps_wsum_kernel<<<g, kThreads, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

int bz_flag_barrier(const BzFlagBarrierArgs* args, cudaStream_t stream) {
  if (args->world < 1 || args->world > BZ_MAXW) return (int)cudaErrorInvalidValue;
  flag_barrier_kernel<<<1, 32, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

int bz_gram_exchange(const BzGramExchangeArgs* args, cudaStream_t stream) {
  if (args->world < 1 || args->world > BZ_MAXW || args->n < 1 || args->n > BZ_MAXN + 16)
    return (int)cudaErrorInvalidValue;
  if (args->slots_mc != nullptr && args->world > 1)
    gram_exchange_nvls_kernel<<<1, 1024, 0, stream>>>(*args);
  else
    gram_exchange_kernel<<<1, 256, 0, stream>>>(*args);
  return (int)cudaGetLastError();
}

int bz_fused_ps_cw(const BzFusedPsArgs* args, int sm_count, cudaStream_t stream) {
  const BzFusedPsArgs a = normalised(*args);
  const int nt = a.n + a.virt.count;
  if (a.n < 1 || nt > BZ_MAXN || a.world < 1 || a.world > BZ_MAXW || a.rank < 0 || a.rank >= a.world)
    return (int)cudaErrorInvalidValue;
  // the fused kernel is vector-only: arenas are 16-byte aligned and padded by construction
  if (!ranges_ok(a)) return (int)cudaErrorInvalidValue;
  for (int i = 0; i < a.n; ++i)
    if (((uintptr_t)a.rows.p[i] % 16) != 0) return (int)cudaErrorInvalidValue;
  for (int p = 0; p < a.world; ++p)
    if (((uintptr_t)a.agg[p] % 16) != 0) return (int)cudaErrorInvalidValue;
  for (int r = 0; r < a.upd.count; ++r)
    if (((uintptr_t)a.upd.param[r] % 16) != 0 ||
        (a.upd.mom[r] && ((uintptr_t)a.upd.mom[r] % 16) != 0))
      return (int)cudaErrorInvalidValue;
  switch (a.mode) {
    case BZ_CW_MEDIAN: return dispatch_mode<BZ_CW_MEDIAN>(a, sm_count, stream, false);
    case BZ_CW_TRMEAN: return dispatch_mode<BZ_CW_TRMEAN>(a, sm_count, stream, false);
    case BZ_CW_MEAMED: return dispatch_mode<BZ_CW_MEAMED>(a, sm_count, stream, false);
    case BZ_CW_MEAN: return dispatch_mode<BZ_CW_MEAN>(a, sm_count, stream, false);
    default: return (int)cudaErrorInvalidValue;
  }
}
