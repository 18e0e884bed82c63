// Coordinate-wise robust aggregation kernels for sm_100a.
//
// One streaming pass over the (n, d) gradient matrix: every thread owns V
// consecutive coordinates, pulls the n values of each coordinate straight from
// the n row buffers (which may be peer-GPU HBM mapped over NVLink), runs a
// register-resident bitonic selection network and writes ONE output value.
//   bytes moved = n*d*4 (read) + d*4 (write)  -> HBM / NVLink bound.
//
// Behavioural parity targets (semantics only, no code shared):
//   median        reference aggregators/coordinate_wise/median.py:102-106   (lower median)
//   trimmed mean  reference aggregators/coordinate_wise/trimmed_mean.py:110-115
//   mean of meds  reference aggregators/coordinate_wise/mean_of_medians.py:71-81
// Attack folding: per-row scale (SignFlip, attacks/sign_flip.py:47-52) and
// virtual rows mu + z*sigma / scale*mean (Little little.py:113-131, Empire
// empire.py:85-92) are synthesised in registers.
// Optional epilogue: SGD(+momentum) update of local replicas (examples/ps/nodes.py:123-125).
#include "cw_core.cuh"

namespace {
using namespace bzcw;

template <int NP, int V, int MODE>
__global__ void __launch_bounds__(kThreads) cw_select_kernel(const __grid_constant__ BzCwArgs a) {
  const long long nvec = a.len / V;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
    const long long base = a.off + u * V;
    float res[V];
    cw_unit<NP, V, MODE>(a.rows, a.scales, a.n, a.virt, a.f, base, res);
    if (a.out != nullptr) {
      if constexpr (V == 4) {
        stg_stream4(a.out + base, make_float4(res[0], res[1], res[2], res[3]));
      } else {
#pragma unroll
        for (int c = 0; c < V; ++c) a.out[base + c] = res[c];
      }
    }
    if (a.upd.count > 0) sgd_apply<V>(a.upd, base, res);
  }
}

// Software-pipelined variant: kStages-deep cp.async ring of thread-private tiles.
constexpr int kStages = 3;

template <int NP, int V, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) cw_select_staged_kernel(const __grid_constant__ BzCwArgs a) {
  extern __shared__ __align__(16) float stage_mem[];
  const int n = a.n;
  const size_t stage_elems = (size_t)n * THREADS * V;
  const long long nvec = a.len / V;
  const long long stride = (long long)gridDim.x * THREADS;
  const long long u0 = (long long)blockIdx.x * THREADS + threadIdx.x;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    const long long u = u0 + s * stride;
    if (u < nvec) cw_stage_issue<NP, V>(stage_mem + s * stage_elems, THREADS, a.rows, n, a.off + u * V);
    cp_async_commit();
  }
  int slot = 0;
  for (long long u = u0; u < nvec; u += stride) {
    {
      const long long un = u + (kStages - 1) * stride;
      int sn = slot + kStages - 1;
      if (sn >= kStages) sn -= kStages;
      if (un < nvec) cw_stage_issue<NP, V>(stage_mem + sn * stage_elems, THREADS, a.rows, n, a.off + un * V);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();
    float v[V][NP];
    cw_stage_read<NP, V>(stage_mem + slot * stage_elems, THREADS, a.scales, n, v);
    float res[V];
    cw_finish<NP, V, MODE>(v, n, a.virt, a.f, res);
    const long long base = a.off + u * V;
    if (a.out != nullptr) {
      if constexpr (V == 4) {
        stg_stream4(a.out + base, make_float4(res[0], res[1], res[2], res[3]));
      } else {
#pragma unroll
        for (int c = 0; c < V; ++c) a.out[base + c] = res[c];
      }
    }
    if (a.upd.count > 0) sgd_apply<V>(a.upd, base, res);
    if (++slot == kStages) slot = 0;
  }
  cp_async_wait<0>();
}

// Warp-tiled variant for MANY rows (NP >= 32), where the thread-private pipeline above spends more alu
// issue slots on feeding the network than on the network itself: with one coordinate per thread every
// row costs a 4-byte cp.async, a 64-bit address add, a predicate and a pointer fetch PER COORDINATE, and
// the -inf / +inf padding is re-selected slot by slot for every coordinate (SASS of the n = 64 median:
// 811 FMNMX of 2530 alu-pipe instructions, profiles/cw_select.md section 4).  Here a WARP owns a tile of
// 32 consecutive coordinates x NP rows in shared memory:
//   * loads: lane l copies the 16-byte chunk (l % 8) of row 4 j + l / 8 -- one warp-wide cp.async
//     instruction moves four 128-byte row segments, so a tile takes NP / 4 instructions instead of NP,
//     each with ONE address computation for four coordinates;
//   * the pad rows n..NP-1 of the tile are written once, before the loop (cp.async never touches them),
//     so the read loop is NP unconditional LDS with immediate offsets and the network runs without the
//     per-slot padding selects (cw_pick<PREPAD>);
//   * NaN canonicalisation is one FMNMX (canon_min) instead of FSETP + FSEL;
//   * lane l then owns coordinate l of the tile: row i of the tile is 32 consecutive floats, so both the
//     chunked writes and the column reads are bank-conflict free.
// The tile is warp-private: cp.async.wait_group + __syncwarp() order the copies of the other lanes before
// the reads, a second __syncwarp() orders the reads before the slot is refilled; no block barrier.
template <int NP, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) cw_select_tiled_kernel(const __grid_constant__ BzCwArgs a) {
  extern __shared__ __align__(16) float stage_mem[];
  constexpr int kWarps = THREADS / 32;
  constexpr int kTile = 32;                                      // coordinates per warp tile
  constexpr size_t kStageElems = (size_t)kWarps * NP * kTile;
  const int n = a.n;
  const int nt = n + a.virt.count;
  const int apad = NP / 2 - 1 - (nt - 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* const wtile = stage_mem + (size_t)warp * NP * kTile;    // + stage * kStageElems
  // rows n..nt-1 are the synthesised rows (written into registers by cw_finish), nt..NP-1 the padding
  for (int s = 0; s < kStages; ++s)
    for (int i = n; i < NP; ++i)
      wtile[s * kStageElems + (size_t)i * kTile + lane] = (i < nt) ? 0.f : ((i - nt < apad) ? -kInf : kInf);
  __syncwarp();
  const long long ntiles = a.len / kTile;                        // the launcher passes whole tiles only
  const long long wstride = (long long)gridDim.x * kWarps;
  const long long w0 = (long long)blockIdx.x * kWarps + warp;
  const int sub = lane >> 3;                                     // row within a group of four
  const int chunk = (lane & 7) * 4;                              // first coordinate of this lane's chunk
  auto issue = [&](int s, long long t) {
    const long long base = a.off + t * kTile + chunk;
    float* dst = wtile + s * kStageElems + chunk;
#pragma unroll
    for (int j = 0; j < NP / 4; ++j) {
      const int r = j * 4 + sub;
      if (r < n) cp_async<16>(dst + r * kTile, a.rows.p[r] + base);
    }
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    const long long t = w0 + s * wstride;
    if (t < ntiles) issue(s, t);
    cp_async_commit();
  }
  int slot = 0;
  for (long long t = w0; t < ntiles; t += wstride) {
    {
      const long long tn = t + (kStages - 1) * wstride;
      int sn = slot + kStages - 1;
      if (sn >= kStages) sn -= kStages;
      if (tn < ntiles) issue(sn, tn);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();
    __syncwarp();                                                // every lane's copies of this slot have landed
    float v[1][NP];
    const float* src = wtile + slot * kStageElems + lane;
#pragma unroll
    for (int i = 0; i < NP; ++i) v[0][i] = canon_min(src[i * kTile] * a.scales.s[i]);
    __syncwarp();                                                // all reads done before the slot is refilled
    float res[1];
    cw_finish<NP, 1, MODE, true>(v, n, a.virt, a.f, res);
    const long long base = a.off + t * kTile + lane;
    if (a.out != nullptr) a.out[base] = res[0];
    if (a.upd.count > 0) sgd_apply<1>(a.upd, base, res);
    if (++slot == kStages) slot = 0;
  }
  cp_async_wait<0>();
}

#ifdef BZ_HOST_EMU
}  // namespace  (the emulator build takes the kernels only; launches are emulated by the harness)
#else
template <int NP, int V, int MODE>
int launch_direct(const BzCwArgs& a, int sm_count, cudaStream_t stream) {
  const long long nvec = a.len / V;
  if (nvec <= 0) return 0;
  long long blocks = (nvec + kThreads - 1) / kThreads;
  // persistent grid: exactly the number of CTAs the device can keep resident (148 SMs x the
  // occupancy of this instantiation), so the grid-stride loop has no partial last wave.
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, cw_select_kernel<NP, V, MODE>, kThreads, 0) !=
            cudaSuccess || o < 1)
      o = 2;
    occ = o;
  }
  const long long cap = (long long)sm_count * occ;
  if (blocks > cap) blocks = cap;
  cw_select_kernel<NP, V, MODE><<<(unsigned)blocks, kThreads, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

template <int NP, int V, int MODE>
int launch_staged(const BzCwArgs& a, int sm_count, cudaStream_t stream) {
  constexpr int THREADS = (NP >= 128) ? 128 : 256;
  const long long nvec = a.len / V;
  if (nvec <= 0) return 0;
  const size_t smem = (size_t)kStages * a.n * THREADS * V * sizeof(float);
  if (smem > 220 * 1024) return launch_direct<NP, V, MODE>(a, sm_count, stream);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cw_select_staged_kernel<NP, V, MODE, THREADS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cw_select_staged_kernel<NP, V, MODE, THREADS>,
                                                    THREADS, smem) != cudaSuccess || occ < 1)
    occ = 1;
  long long blocks = (nvec + THREADS - 1) / THREADS;
  const long long cap = (long long)sm_count * occ;
  if (blocks > cap) blocks = cap;
  cw_select_staged_kernel<NP, V, MODE, THREADS><<<(unsigned)blocks, THREADS, smem, stream>>>(a);
  return (int)cudaGetLastError();
}

// Whole 32-coordinate tiles through the warp-tiled kernel; returns the number of coordinates it took
// (0 = not applicable here), or a negative cudaError.
template <int NP, int MODE>
long long launch_tiled(const BzCwArgs& a, int sm_count, cudaStream_t stream) {
  if constexpr (NP < 32 || MODE == BZ_CW_MEAN || (NP >= 128 && MODE == BZ_CW_MEAMED)) {
    return 0;       // (mean-of-medians at 128 rows keeps three 128-value arrays live: it spills here)
  } else {
    constexpr int THREADS = (NP >= 128) ? 128 : 256;
    constexpr int kWarps = THREADS / 32;
    const long long main_len = a.len - (a.len % 32);
    if (main_len <= 0) return 0;
    const size_t smem = (size_t)kStages * kWarps * NP * 32 * sizeof(float);
    if (smem > 220 * 1024) return 0;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(cw_select_tiled_kernel<NP, MODE, THREADS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return -(long long)e;
      configured = true;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cw_select_tiled_kernel<NP, MODE, THREADS>, THREADS,
                                                      smem) != cudaSuccess || occ < 1)
      occ = 1;
    BzCwArgs b = a;
    b.len = main_len;
    for (int i = a.n; i < BZ_MAXN; ++i) b.scales.s[i] = 1.f;     // pad / synthesised rows are not scaled
    const long long ntiles = main_len / 32;
    long long blocks = (ntiles + kWarps - 1) / kWarps;
    const long long cap = (long long)sm_count * occ;
    if (blocks > cap) blocks = cap;
    cw_select_tiled_kernel<NP, MODE, THREADS><<<(unsigned)blocks, THREADS, smem, stream>>>(b);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? main_len : -(long long)e;
  }
}

template <int NP, int V, int MODE>
int launch_one(const BzCwArgs& a, int sm_count, cudaStream_t stream) {
  // auto: the direct kernel is at the HBM roofline for the pruned median network up to 8 rows and
  // for the plain mean; everything with a longer ALU phase per tile is latency bound without the
  // cp.async pipeline (profiles/cw_select.md).
  bool staged = (MODE == BZ_CW_TRMEAN || MODE == BZ_CW_MEAMED || (MODE == BZ_CW_MEDIAN && NP >= 16));
  if (a.impl == 1) staged = false;
  if (a.impl == 2) staged = true;
  // (impl 3 = warp-tiled, see launch_np: where that kernel does not apply the automatic choice stands)
  // short vectors: one tile per thread, nothing to pipeline
  if (a.impl == 0 && a.len / V < (long long)sm_count * kThreads * 2) staged = false;
  return staged ? launch_staged<NP, V, MODE>(a, sm_count, stream)
                : launch_direct<NP, V, MODE>(a, sm_count, stream);
}

template <int NP, int MODE>
int launch_np(const BzCwArgs& a, bool vec_ok, int sm_count, cudaStream_t stream) {
  constexpr int VMAX = (NP <= 16) ? 4 : (NP == 32 ? 2 : 1);
  if (a.impl == 3 && vec_ok) {
    // warp-tiled kernel on the whole tiles, the (< 32 coordinate) tail through the scalar kernel
    const long long took = launch_tiled<NP, MODE>(a, sm_count, stream);
    if (took < 0) return (int)(-took);
    if (took > 0) {
      if (took == a.len) return 0;
      BzCwArgs b = a;
      b.impl = 0;
      b.off = a.off + took;
      b.len = a.len - took;
      return launch_one<NP, 1, MODE>(b, sm_count, stream);
    }
  }
  if constexpr (VMAX > 1) {
    if (vec_ok) {
      const long long main_len = a.len - (a.len % VMAX);
      BzCwArgs b = a;
      b.len = main_len;
      int e = launch_one<NP, VMAX, MODE>(b, sm_count, stream);
      if (e) return e;
      if (main_len < a.len) {
        b.off = a.off + main_len;
        b.len = a.len - main_len;
        e = launch_one<NP, 1, MODE>(b, sm_count, stream);
      }
      return e;
    }
  }
  return launch_one<NP, 1, MODE>(a, sm_count, stream);
}

template <int MODE>
int launch_mode(const BzCwArgs& a, bool vec_ok, int sm_count, cudaStream_t stream) {
  const int nt = a.n + a.virt.count;
  if (nt <= 2) return launch_np<2, MODE>(a, vec_ok, sm_count, stream);
  if (nt <= 4) return launch_np<4, MODE>(a, vec_ok, sm_count, stream);
  if (nt <= 8) return launch_np<8, MODE>(a, vec_ok, sm_count, stream);
  if (nt <= 16) return launch_np<16, MODE>(a, vec_ok, sm_count, stream);
  if (nt <= 32) return launch_np<32, MODE>(a, vec_ok, sm_count, stream);
  if (nt <= 64) return launch_np<64, MODE>(a, vec_ok, sm_count, stream);
  return launch_np<128, MODE>(a, vec_ok, sm_count, stream);
}

}  // namespace

int bz_cw_select(const BzCwArgs* args, int sm_count, cudaStream_t stream) {
  const BzCwArgs& a = *args;
  const int nt = a.n + a.virt.count;
  if (nt < 1 || nt > BZ_MAXN || a.n < 1) return (int)cudaErrorInvalidValue;
  if (a.upd.count < 0 || a.upd.count > BZ_MAXR) return (int)cudaErrorInvalidValue;
  // Vector path needs 16-byte alignment of every stream touched.
  bool vec_ok = (a.off % 4) == 0;
  for (int i = 0; i < a.n && vec_ok; ++i) vec_ok = ((uintptr_t)a.rows.p[i] % 16) == 0;
  if (a.out) vec_ok = vec_ok && ((uintptr_t)a.out % 16) == 0;
  for (int r = 0; r < a.upd.count && vec_ok; ++r) {
    vec_ok = ((uintptr_t)a.upd.param[r] % 16) == 0 &&
             (a.upd.mom[r] == nullptr || ((uintptr_t)a.upd.mom[r] % 16) == 0);
  }
  switch (a.mode) {
    case BZ_CW_MEDIAN: return launch_mode<BZ_CW_MEDIAN>(a, vec_ok, sm_count, stream);
    case BZ_CW_TRMEAN: return launch_mode<BZ_CW_TRMEAN>(a, vec_ok, sm_count, stream);
    case BZ_CW_MEAMED: return launch_mode<BZ_CW_MEAMED>(a, vec_ok, sm_count, stream);
    case BZ_CW_MEAN: return launch_mode<BZ_CW_MEAN>(a, vec_ok, sm_count, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
#endif  // BZ_HOST_EMU
