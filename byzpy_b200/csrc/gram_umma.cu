// Gram matrix  G = (S X)(S X)^T  on the 5th-generation tensor cores (tcgen05 / UMMA),
// with TMEM accumulators.  sm_100a only.
//
// Shape: M = N = n <= 128 rows (padded to M = 128, N = round_up(n, 16)), K = d (huge):
// a split-K skinny GEMM whose cost is ONE read of the n*d matrix.  Each persistent CTA owns a
// strided set of 32-column K tiles and accumulates all of them into the SAME TMEM
// accumulators; per-CTA partial results are reduced in fp64 by a second tiny kernel
// (deterministic two-stage split-K, no atomics).
//
// Block-diagonal packing: the MMA atom is always M = N = 128.  For n < 128 the 128 operand rows hold
// B = 128 / n_pad different 32-column blocks of the SAME n rows (n_pad = n rounded up to 16/32/64/128),
// i.e. one tile covers 32*B columns.  The 128 x 128 product then contains the B partial Grams on
// its diagonal n_pad x n_pad blocks (off-diagonal blocks are cross terms and are ignored), so the
// bytes moved per MMA are the same 16 KB for every n and the kernel runs at the n = 128 rate.
//
// Precision (SURVEY 7.1 "precision hazard"): kind::tf32 keeps 10 mantissa bits, which is not
// enough for G_ii + G_jj - 2 G_ij when gradients are close.  Every fp32 value is split
// x = hi + lo with hi = rna_tf32(x); the kernel accumulates  HH = hi hi^T  and  HL = hi lo^T
// in two TMEM accumulators, and the reduction forms  G = HH + HL + HL^T  (the dropped lo lo^T
// term is 2^-22 relative).  Two MMAs per k-step instead of three because A == B == X.
//
// Warp roles (288 threads):
//   warp 0      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases the
//               operand stage / signals the epilogue;
//   warps 1..8  loaders + converters: 16-byte cp.async straight from the row pointer table (rows may
//               live in peer HBM) into a THREAD-PRIVATE slot of a 6-deep shared-memory ring (5 raw
//               tiles = 80 KB in flight per SM; only cp.async.wait_group, no block barrier); the owner
//               reads its chunk back, splits hi/lo and writes the two K-major SWIZZLE_128B operand
//               tiles (conflict-free), fence.proxy.async, arrive on the stage's mbarrier.
//               Warps 1..4 are also the epilogue (tcgen05.ld TMEM -> registers -> per-CTA partial).
// Two mbarrier rings: operand full/empty (converters <-> MMA, 3 stages) and accumulator full.
//
// Measured design note (profiles/gram_umma.md): the first version fed the converters with one
// cp.async.bulk (TMA, UBLKCP) per row per 64-column chunk.  Bulk copies are issued through the
// uniform datapath, ~80 cycles each, so 256-byte copies cap a CTA at ~3 B/clk (0.9 TB/s chip-wide,
// ncu: dram 11 %, tensor 10 %).  A pointer-table of rows cannot use one 2-D tensor map, so the
// loads moved to the LSU: first LDG with a 3-tile register prefetch (0.40 of the HBM roofline),
// then the cp.async ring above (0.52; tensor pipe 41 % busy with two TF32 MMAs per k-step).
#include "api.h"
#include "gram_umma.h"

namespace {

constexpr int kConvWarps = 8;                 // converter warps (also: first 4 = epilogue)
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kProdWarp = 1 + kConvWarps;     // warp 9: TMA producer (idle in the cp.async variant)
constexpr int kThreads = 32 + kConvThreads + 32;   // warp 0 = TMEM allocator + MMA issuer
constexpr int kOpCols = 32;                   // columns per operand tile (128 B swizzle row)
constexpr int kRows = 128;                    // padded M
constexpr int kOpStages = 3;
constexpr int kOpTileBytes = kRows * kOpCols * 4;             // 16 KB (hi) ; lo follows
constexpr int kOpStageBytes = 2 * kOpTileBytes;               // 32 KB
constexpr int kRawStages = 6;                 // cp.async ring of raw fp32 tiles: 5 tiles (80 KB) in flight per SM
constexpr int kChunksPerThread = kRows * 8 / kConvThreads;    // 16-byte chunks per thread per tile (4)
constexpr int kRawStageBytes = kChunksPerThread * kConvThreads * 16;   // 16 KB
constexpr int kTmemCols = 256;                // D_hh at column 0, D_hl at column 128
constexpr int kBarBytes = 256;                // mbarriers + TMEM slot (keeps the raw ring 128-byte aligned)
constexpr int kSmemBytes = 1024 /*align slack*/ + kOpStages * kOpStageBytes + kBarBytes +
                           kRawStages * kRawStageBytes;
constexpr unsigned long long kWaitBudgetCycles = 4000000000ull;  // ~2 s: trap instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWaitBudgetCycles) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA: 2-D tile global -> shared, completion (bytes) on an mbarrier.  SASS: UTMALDG.2D
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO (unused for swizzled K-major, set to 1) | SBO = 1024 B between
//   8-row groups | layout type 2 (SWIZZLE_128B) in bits [61, 64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // bits [0, 14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (16 B units)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = n_pad.
__device__ __forceinline__ uint32_t make_idesc(int n_pad) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_pad >> 3) << 17) |
         ((uint32_t)(kRows >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool kTma>
__device__ __forceinline__ void gram_umma_body(const BzGramUmmaArgs& a, const BzGramTmaMaps* tm) {
  extern __shared__ uint8_t smem_raw_[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* op_base = (uint8_t*)(((uintptr_t)smem_raw_ + 1023) & ~(uintptr_t)1023);   // kOpStages x (hi | lo)
  uint64_t* bars = (uint64_t*)(op_base + kOpStages * kOpStageBytes);
  uint64_t* op_full = bars;                     // [kOpStages]  converters -> MMA
  uint64_t* op_empty = bars + kOpStages;        // [kOpStages]  MMA (tcgen05.commit) -> converters
  uint64_t* acc_full = op_empty + kOpStages;    // [1]          MMA -> epilogue
  uint64_t* raw_full = acc_full + 1;            // [kRawStages] TMA (complete_tx) -> converters
  uint64_t* raw_empty = raw_full + kRawStages;  // [kRawStages] converters -> TMA producer
  uint32_t* tmem_slot = (uint32_t*)(raw_empty + kRawStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n;
  const int n_pad = a.n_pad;                                      // 16 / 32 / 64 / 128
  const int nblk = kRows / n_pad;                                 // column blocks packed per tile
  const int tile_cols = kOpCols * nblk;
  const long long ntiles = a.len / tile_cols;                     // full tiles only
  long long my_tiles = 0;
  if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;

  // ---- one-time setup ------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < kOpStages; ++s) {
      mbar_init(smem_u32(&op_full[s]), kConvWarps);
      mbar_init(smem_u32(&op_empty[s]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(smem_u32(&raw_full[s]), 1);
      mbar_init(smem_u32(&raw_empty[s]), kConvWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows >= n of every operand tile stay zero for the whole kernel
  for (int i = threadIdx.x; i < kOpStages * kOpStageBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(op_base)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ====================================== MMA issuer ====================================
    const uint32_t idesc = make_idesc(kRows);
    const uint32_t d_hh = tmem_base, d_hl = tmem_base + 128;
    for (long long t = 0; t < my_tiles; ++t) {
      const int os = (int)(t % kOpStages);
      mbar_wait(smem_u32(&op_full[os]), (uint32_t)((t / kOpStages) & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t hi = smem_u32(op_base + os * kOpStageBytes);
        const uint32_t lo = hi + kOpTileBytes;
#pragma unroll
        for (int kk = 0; kk < kOpCols / 8; ++kk) {   // K = 8 tf32 (32 bytes) per instruction
          const uint64_t dh = make_desc(hi + kk * 32);
          const uint64_t dl = make_desc(lo + kk * 32);
          const uint32_t acc = (t > 0 || kk > 0) ? 1u : 0u;
          umma_tf32(d_hh, dh, dh, idesc, acc);
          umma_tf32(d_hl, dh, dl, idesc, acc);
        }
        umma_commit(smem_u32(&op_empty[os]));         // frees the operand stage when the MMAs retire
        if (t == my_tiles - 1) umma_commit(smem_u32(acc_full));
      }
      __syncwarp();
    }
  } else if (warp == kProdWarp) {
    // ================================ TMA producer (one thread) ===========================
    if constexpr (kTma) {
      if (lane == 0) {
        uint8_t* raw_base = op_base + kOpStages * kOpStageBytes + kBarBytes;
        const uint32_t row_bytes = (uint32_t)tile_cols * 4u;
        uint32_t tile_bytes = 0;
        for (int k = 0; k < tm->nseg; ++k) tile_bytes += (uint32_t)tm->seg_rows[k] * row_bytes;
        for (long long t = 0; t < my_tiles; ++t) {
          const int rs = (int)(t % kRawStages);
          mbar_wait(smem_u32(&raw_empty[rs]), (uint32_t)(((t / kRawStages) & 1) ^ 1));
          const long long col = a.off + ((long long)blockIdx.x + t * gridDim.x) * tile_cols;
          const uint32_t bar = smem_u32(&raw_full[rs]);
          mbar_expect_tx(bar, tile_bytes);
          uint8_t* dst = raw_base + (size_t)rs * kRawStageBytes;
          for (int k = 0; k < tm->nseg; ++k)
            tma_load_2d(smem_u32(dst + (size_t)tm->seg_row0[k] * row_bytes), &tm->maps[k], (int)col, 0, bar);
        }
      }
    }
  } else {
    // ========================= loaders / converters / epilogue ============================
    const int ct = threadIdx.x - 32;
    uint8_t* raw_base = op_base + kOpStages * kOpStageBytes + kBarBytes;
    if constexpr (kTma) {
      // The raw tile of a stage is [n rows][tile_cols] fp32, row pitch tile_cols * 4 bytes, written by
      // the TMA.  Chunk q of a thread: data row q / (8 nblk), 16-byte chunk c = q % (8 nblk) of that row
      // (consecutive threads read consecutive chunks: conflict-free); it belongs to column block c / 8
      // and lands in operand row blk * n_pad + drow.
      const int cpr = 8 * nblk;                        // 16-byte chunks per data row
      int rows_[kChunksPerThread], cs_[kChunksPerThread], raw_off[kChunksPerThread];
#pragma unroll
      for (int j = 0; j < kChunksPerThread; ++j) {
        const int q = j * kConvThreads + ct;
        const int drow = q / cpr, c = q % cpr;
        rows_[j] = (drow < n) ? (c >> 3) * n_pad + drow : -1;
        cs_[j] = c & 7;
        raw_off[j] = drow * (tile_cols * 4) + c * 16;
      }
      for (long long t = 0; t < my_tiles; ++t) {
        const int os = (int)(t % kOpStages);
        const int rs = (int)(t % kRawStages);
        mbar_wait(smem_u32(&raw_full[rs]), (uint32_t)((t / kRawStages) & 1));
        mbar_wait(smem_u32(&op_empty[os]), (uint32_t)(((t / kOpStages) & 1) ^ 1));
        const uint8_t* raw = raw_base + (size_t)rs * kRawStageBytes;
        uint8_t* hi = op_base + os * kOpStageBytes;
        uint8_t* lo = hi + kOpTileBytes;
#pragma unroll
        for (int j = 0; j < kChunksPerThread; ++j) {
          if (rows_[j] >= 0) {
            const float4 x = *reinterpret_cast<const float4*>(raw + raw_off[j]);
            const float xs[4] = {x.x, x.y, x.z, x.w};      // row scales are applied in the reduce
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hh[e] = to_tf32(xs[e]);
              ll[e] = __float_as_uint(xs[e] - __uint_as_float(hh[e]));
            }
            const int row = rows_[j];
            const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((cs_[j] ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
        fence_proxy_async();                         // generic-proxy writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&raw_empty[rs]));      // this warp is done reading the raw stage
          mbar_arrive(smem_u32(&op_full[os]));        // one arrival per converter warp
        }
      }
    } else {
    // Each thread owns kChunksPerThread 16-byte chunks of every 128 x 32 tile: chunk q covers
    // row q / 8, 16-byte column group q % 8.  The raw fp32 chunks travel global (local or peer HBM)
    // -> shared memory with cp.async into a THREAD-PRIVATE slot of a kRawStages-deep ring, so the
    // bytes in flight are bounded by shared memory (80 KB / SM) instead of registers and nobody
    // else ever reads the slot (cp.async.wait_group is the only synchronisation); the owner then
    // splits hi / lo and writes the swizzled operand tiles.
    int rows_[kChunksPerThread], cs_[kChunksPerThread];
    const float* src_[kChunksPerThread];
#pragma unroll
    for (int j = 0; j < kChunksPerThread; ++j) {
      const int q = j * kConvThreads + ct;
      rows_[j] = q >> 3;                               // tile row = block * n_pad + data row
      cs_[j] = q & 7;
      const int blk = rows_[j] / n_pad, drow = rows_[j] % n_pad;
      src_[j] = (drow < n) ? a.rows.p[drow] + a.off + blk * kOpCols + cs_[j] * 4 : nullptr;
    }
    auto raw_slot = [&](int stage, int j) -> uint8_t* {
      return raw_base + (size_t)stage * kRawStageBytes + ((size_t)j * kConvThreads + ct) * 16;
    };
    auto issue = [&](long long t) {
      const long long col = ((long long)blockIdx.x + t * gridDim.x) * tile_cols;
      const int stage = (int)(t % kRawStages);
#pragma unroll
      for (int j = 0; j < kChunksPerThread; ++j) {
        if (src_[j] != nullptr) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(raw_slot(stage, j))),
                       "l"(src_[j] + col)
                       : "memory");
        }
      }
    };
#pragma unroll
    for (int pf = 0; pf < kRawStages - 1; ++pf) {
      if (pf < my_tiles) issue(pf);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (long long t = 0; t < my_tiles; ++t) {
      if (t + kRawStages - 1 < my_tiles) issue(t + kRawStages - 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kRawStages - 1) : "memory");
      const int os = (int)(t % kOpStages);
      const int rs = (int)(t % kRawStages);
      mbar_wait(smem_u32(&op_empty[os]), (uint32_t)(((t / kOpStages) & 1) ^ 1));
      uint8_t* hi = op_base + os * kOpStageBytes;
      uint8_t* lo = hi + kOpTileBytes;
#pragma unroll
      for (int j = 0; j < kChunksPerThread; ++j) {
        if (src_[j] != nullptr) {
          const float4 x = *reinterpret_cast<const float4*>(raw_slot(rs, j));
          const float xs[4] = {x.x, x.y, x.z, x.w};      // row scales are applied in the reduce
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hh[e] = to_tf32(xs[e]);
            ll[e] = __float_as_uint(xs[e] - __uint_as_float(hh[e]));
          }
          const int row = rows_[j];
          const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((cs_[j] ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
      }
      fence_proxy_async();                         // generic-proxy writes -> async proxy (UMMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&op_full[os]));   // one arrival per converter warp
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // ---- epilogue: TMEM -> registers -> per-CTA partials (first 4 converter warps) -------
    if (my_tiles > 0 && warp <= 4) {
      mbar_wait(smem_u32(acc_full), 0);
      tc_fence_after();
      const int quad = warp & 3;                       // TMEM lane partition of this warp
      const int trow = quad * 32 + lane;               // accumulator row owned by this thread
      const int my_blk = trow / n_pad, drow = trow % n_pad;
      // tcgen05.ld is warp-collective: the TMEM address must be warp-uniform.  A warp's 32 rows
      // span max(1, 32 / n_pad) diagonal blocks; load each block's columns with a uniform address
      // and let the lanes that belong to that block keep the data.
      const int blk_first = (quad * 32) / n_pad;
      const int blk_last = (quad * 32 + 31) / n_pad;
      for (int blk = blk_first; blk <= blk_last; ++blk) {
        // partial layout: [cta][block][HH | HL][n][n]
        float* PA = a.partials + ((size_t)blockIdx.x * nblk + blk) * 2 * n * n;
        float* PB = PA + (size_t)n * n;
        for (int c0 = 0; c0 < n_pad; c0 += 16) {
          uint32_t va[16], vb[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(blk * n_pad + c0);
          tmem_ld16(taddr, va);
          tmem_ld16(taddr + 128, vb);
          if (blk == my_blk && drow < n) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c0 + j < n) {
                PA[drow * n + c0 + j] = __uint_as_float(va[j]);
                PB[drow * n + c0 + j] = __uint_as_float(vb[j]);
              }
            }
          }
        }
      }
    }
  }
  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

__global__ void __launch_bounds__(kThreads, 1) gram_umma_kernel(const __grid_constant__ BzGramUmmaArgs a) {
  gram_umma_body<false>(a, nullptr);
}
__global__ void __launch_bounds__(kThreads, 1) gram_umma_tma_kernel(const __grid_constant__ BzGramUmmaArgs a,
                                                                   const __grid_constant__ BzGramTmaMaps tm) {
  gram_umma_body<true>(a, &tm);
}

// G_ij = s_i s_j * ( sum_c HH_c[i][j] + HL_c[i][j] + HL_c[j][i] ) + G_tail[i][j]
__global__ void gram_umma_reduce_kernel(const float* __restrict__ partials, int num_partials, int n,
                                        ScaleTable scales, const double* __restrict__ tail64,
                                        float* __restrict__ G, double* __restrict__ G64) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  const int i = t / n, j = t % n;
  double s = 0.0;
  for (int c = 0; c < num_partials; ++c) {
    const float* PA = partials + (size_t)c * 2 * n * n;
    const float* PB = PA + (size_t)n * n;
    s += (double)PA[i * n + j] + (double)PB[i * n + j] + (double)PB[j * n + i];
  }
  s *= (double)scales.s[i] * (double)scales.s[j];
  if (tail64) s += tail64[t];
  G[t] = (float)s;
  if (G64) G64[t] = s;
}

}  // namespace

int bz_gram_umma_npad(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128)); }

int bz_gram_umma_tile_cols(int n) { return kOpCols * (kRows / bz_gram_umma_npad(n)); }

int bz_gram_umma_grid(int n, long long len, int sm_count) {
  const long long ntiles = len / bz_gram_umma_tile_cols(n);
  if (ntiles <= 0) return 0;
  return (int)(ntiles < sm_count ? ntiles : sm_count);
}

int bz_gram_umma_partials(int n, int grid) { return grid * (kRows / bz_gram_umma_npad(n)); }

static int gram_umma_launch(const BzGramUmmaArgs* args, const BzGramTmaMaps* maps, int sm_count,
                            cudaStream_t stream);

int bz_gram_umma(const BzGramUmmaArgs* args, int sm_count, cudaStream_t stream) {
  return gram_umma_launch(args, nullptr, sm_count, stream);
}

int bz_gram_umma_tma(const BzGramUmmaArgs* args, const BzGramTmaMaps* maps, int sm_count, cudaStream_t stream) {
  if (maps == nullptr || maps->nseg < 1 || maps->nseg > BZ_GRAM_MAXSEG) return (int)cudaErrorInvalidValue;
  int covered = 0;
  for (int k = 0; k < maps->nseg; ++k) {
    if (maps->seg_row0[k] != covered || maps->seg_rows[k] < 1) return (int)cudaErrorInvalidValue;
    covered += maps->seg_rows[k];
  }
  if (covered != args->n) return (int)cudaErrorInvalidValue;
  return gram_umma_launch(args, maps, sm_count, stream);
}

static int gram_umma_launch(const BzGramUmmaArgs* args, const BzGramTmaMaps* maps, int sm_count,
                            cudaStream_t stream) {
  const BzGramUmmaArgs& a = *args;
  if (a.n < 1 || a.n > BZ_MAXN) return (int)cudaErrorInvalidValue;
  if ((a.off % 4) != 0) return (int)cudaErrorInvalidValue;
  for (int i = 0; i < a.n; ++i)
    if (((uintptr_t)a.rows.p[i] % 16) != 0) return (int)cudaErrorInvalidValue;
  BzGramUmmaArgs b = *args;
  b.n_pad = bz_gram_umma_npad(a.n);
  const int grid = bz_gram_umma_grid(a.n, a.len, sm_count);
  const int nparts = bz_gram_umma_partials(a.n, grid);
  if (nparts > a.num_partials) return (int)cudaErrorInvalidValue;
  if (grid > 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(gram_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemBytes);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gram_umma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    if (maps != nullptr)
      gram_umma_tma_kernel<<<grid, kThreads, kSmemBytes, stream>>>(b, *maps);
    else
      gram_umma_kernel<<<grid, kThreads, kSmemBytes, stream>>>(b);
    int e = (int)cudaGetLastError();
    if (e) return e;
  }
  const int rt = 128;
  gram_umma_reduce_kernel<<<(a.n * a.n + rt - 1) / rt, rt, 0, stream>>>(
      a.partials, nparts, a.n, a.scales, a.tail64, a.G, a.G64);
  return (int)cudaGetLastError();
}
