// The coordinate-wise CUDA kernels (csrc/cw_select.cu: direct, thread-private cp.async-staged and warp-tiled
// pipelines) executed on the host by the warp-lockstep emulator of cuda_host_emu.h and compared, coordinate by
// coordinate, with the register code applied to plainly gathered values (that code is itself checked against
// std::sort by cw_network_host.cu).  The direct and staged kernels are GPU-validated: reproducing them validates
// the emulator; the warp-tiled kernel has not run on a GPU yet and this is its functional test.
//
//   g++ -std=c++17 -O1 -DBZ_HOST_EMU -Wno-unknown-pragmas -I byzpy_b200/csrc -I tests/native \
//       tests/native/cw_kernels_emu.cpp -o /tmp/cw_kernels_emu && /tmp/cw_kernels_emu
#include <random>

#include "cw_select.cu"

namespace {
__attribute__((aligned(16))) float stage_mem[220 * 1024 / 4];      // the kernels' `extern __shared__ float stage_mem[]`
}

using namespace bzcw;

static long long g_checks = 0, g_fail = 0;

struct Problem {
  int n, nv, f, mode;
  long long off, len, d;
  std::vector<std::vector<float>> rows;
  std::vector<float> scales;
  VirtRows virt;
};

static Problem make_problem(std::mt19937& rng, int n, int nv, int f, int mode, long long off, long long len, bool scaled,
                            bool weird) {
  Problem p;
  p.n = n; p.nv = nv; p.f = f; p.mode = mode; p.off = off; p.len = len; p.d = off + len + 7;
  std::normal_distribution<float> nd(0.f, 1.f);
  p.rows.assign(n, std::vector<float>(p.d + 8));
  for (auto& r : p.rows)
    for (auto& x : r) {
      x = nd(rng);
      if (weird && (rng() % 13) == 0) x = (rng() & 1) ? std::nanf("") : ((rng() & 1) ? INFINITY : -INFINITY);
    }
  p.scales.assign(BZ_MAXN, 7.f);                  // entries beyond n are garbage on purpose
  for (int i = 0; i < n; ++i) p.scales[i] = scaled ? ((i % 3 == 0) ? -1.f : (i % 3 == 1 ? 0.5f : 1.f)) : 1.f;
  p.virt = nv > 0 ? VirtRows{nv, std::max(1, n - 1), 0.6f, -0.9f} : VirtRows{0, 0, 0.f, 0.f};
  return p;
}

template <int NP, int MODE>
static float expected(const Problem& p, long long j) {
  float v[1][NP];
  for (int i = 0; i < NP; ++i) v[0][i] = (i < p.n) ? canon(p.rows[i][j] * p.scales[i]) : 0.f;
  float res[1];
  cw_finish<NP, 1, MODE, false>(v, p.n, p.virt, p.f, res);
  return res[0];
}

static bool same(float a, float b) {
  if (std::isnan(a) || std::isnan(b)) return std::isnan(a) && std::isnan(b);
  return a == b;                                  // bit-identical: same operations in the same order
}

static BzCwArgs make_args(const Problem& p, float* out, bool tiled_scales, float* param, float* mom) {
  BzCwArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int i = 0; i < p.n; ++i) a.rows.p[i] = p.rows[i].data();
  for (int i = 0; i < BZ_MAXN; ++i) a.scales.s[i] = p.scales[i];
  if (tiled_scales)                               // what launch_tiled() does before launching
    for (int i = p.n; i < BZ_MAXN; ++i) a.scales.s[i] = 1.f;
  a.n = p.n; a.virt = p.virt; a.f = p.f; a.mode = p.mode; a.off = p.off; a.len = p.len; a.out = out;
  if (param) {
    a.upd.count = 1; a.upd.param[0] = param; a.upd.mom[0] = mom; a.upd.lr = 0.1f; a.upd.mu = 0.9f; a.upd.wd = 0.01f;
  }
  return a;
}

template <int NP, int MODE>
static void verify(const char* what, const Problem& p, const std::vector<float>& out, const std::vector<float>* param,
                   const std::vector<float>* param0, const std::vector<float>* mom) {
  const long long before = g_fail;
  for (long long j = p.off; j < p.off + p.len; ++j) {
    const float want = expected<NP, MODE>(p, j);
    ++g_checks;
    if (!same(out[j], want)) {
      if (g_fail - before < 4)
        std::printf("MISMATCH %s NP=%d mode=%d n=%d nv=%d f=%d off=%lld len=%lld j=%lld: got %.9g want %.9g\n", what, NP,
                    MODE, p.n, p.nv, p.f, p.off, p.len, j, out[j], want);
      ++g_fail;
    }
    if (param) {
      float g = want + 0.01f * (*param0)[j];
      const float mo = 0.9f * 0.f + g;
      const float pw = (*param0)[j] - 0.1f * mo;
      ++g_checks;
      if (!same((*param)[j], pw) || !same((*mom)[j], mo)) ++g_fail;
    }
  }
  // nothing outside [off, off + len) may be touched
  for (long long j = 0; j < (long long)out.size(); ++j)
    if ((j < p.off || j >= p.off + p.len) && out[j] != -777.f) {
      if (g_fail - before < 4) std::printf("STRAY WRITE %s at %lld\n", what, j);
      ++g_fail;
    }
}

template <int NP, int MODE, int THREADS>
static void run_tiled(std::mt19937& rng, int n, int nv, int f, long long off, long long tiles, unsigned grid, bool scaled,
                      bool weird, bool update) {
  Problem p = make_problem(rng, n, nv, f, MODE, off, tiles * 32, scaled, weird);
  std::vector<float> out(p.d + 8, -777.f), param(p.d + 8), param0, mom(p.d + 8, 0.f);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (auto& x : param) x = nd(rng);
  param0 = param;
  BzCwArgs a = make_args(p, out.data(), true, update ? param.data() : nullptr, update ? mom.data() : nullptr);
  const size_t smem = (size_t)kStages * (THREADS / 32) * NP * 32;
  emu::launch(grid, THREADS, stage_mem, smem, [&] { cw_select_tiled_kernel<NP, MODE, THREADS>(a); });
  verify<NP, MODE>("tiled", p, out, update ? &param : nullptr, update ? &param0 : nullptr, update ? &mom : nullptr);
}

template <int NP, int V, int MODE, int THREADS>
static void run_staged(std::mt19937& rng, int n, int nv, int f, long long off, long long len, unsigned grid, bool scaled) {
  Problem p = make_problem(rng, n, nv, f, MODE, off, len, scaled, false);
  std::vector<float> out(p.d + 8, -777.f);
  BzCwArgs a = make_args(p, out.data(), false, nullptr, nullptr);
  const size_t smem = (size_t)kStages * n * THREADS * V;
  emu::launch(grid, THREADS, stage_mem, smem, [&] { cw_select_staged_kernel<NP, V, MODE, THREADS>(a); });
  verify<NP, MODE>("staged", p, out, nullptr, nullptr, nullptr);
}

template <int NP, int V, int MODE>
static void run_direct(std::mt19937& rng, int n, int nv, int f, long long off, long long len, unsigned grid, bool scaled) {
  Problem p = make_problem(rng, n, nv, f, MODE, off, len, scaled, false);
  std::vector<float> out(p.d + 8, -777.f);
  BzCwArgs a = make_args(p, out.data(), false, nullptr, nullptr);
  emu::launch(grid, kThreads, stage_mem, 0, [&] { cw_select_kernel<NP, V, MODE>(a); });
  verify<NP, MODE>("direct", p, out, nullptr, nullptr, nullptr);
}

template <int NP, int THREADS>
static void tiled_suite(std::mt19937& rng) {
  const int ns[] = {NP / 2 + 1, NP / 2 + 5, NP - 1, NP};
  for (int n : ns) {
    for (int nv : {0, 2}) {
      if (n + nv > NP) continue;
      const int nt = n + nv;
      // tiles: fewer than the pipeline depth, exactly one per warp, and several rounds of slot rotation
      const unsigned grid = 2;
      const long long warps = (long long)grid * (THREADS / 32);
      for (long long tiles : {1ll, warps, warps * 4 + 3}) {
        const bool big = tiles > warps;
        run_tiled<NP, BZ_CW_MEDIAN, THREADS>(rng, n, nv, 0, 64, tiles, grid, big, nv == 0 && big, false);
        run_tiled<NP, BZ_CW_TRMEAN, THREADS>(rng, n, nv, (nt - 1) / 3, 0, tiles, grid, false, false, big && nv == 0);
        if (NP < 128) run_tiled<NP, BZ_CW_MEAMED, THREADS>(rng, n, nv, nt / 4, 32, tiles, grid, big, false, false);
      }
    }
  }
}

static void suite(bool eager) {
  emu::st().eager = eager;
  const long long c0 = g_checks, f0 = g_fail, e0 = emu::st().errors;
  std::mt19937 rng(99);
  // ---- emulator self-check on the GPU-validated kernels
  run_direct<8, 4, BZ_CW_MEDIAN>(rng, 8, 0, 0, 0, 4 * 300, 1, true);
  run_direct<8, 4, BZ_CW_TRMEAN>(rng, 6, 2, 2, 4, 4 * 513, 2, false);
  run_direct<64, 1, BZ_CW_MEDIAN>(rng, 50, 0, 0, 3, 700, 2, false);
  run_staged<8, 4, BZ_CW_TRMEAN, 256>(rng, 8, 0, 2, 0, 4 * 256 * 7 + 4 * 10, 1, true);
  run_staged<16, 4, BZ_CW_MEDIAN, 256>(rng, 11, 2, 0, 8, 4 * 256 * 4, 2, false);
  run_staged<64, 1, BZ_CW_MEDIAN, 256>(rng, 64, 0, 0, 5, 256 * 5 + 17, 1, false);
  run_staged<64, 1, BZ_CW_MEAMED, 256>(rng, 40, 0, 9, 0, 256 * 3, 2, true);
  const long long self = g_checks;
  std::printf("[%s cp.async] self-check (direct + staged kernels): %lld checks, %lld failures\n",
              eager ? "eager" : "deferred", g_checks - c0, g_fail - f0);
  // ---- the warp-tiled kernel
  tiled_suite<32, 256>(rng);
  tiled_suite<64, 256>(rng);
  tiled_suite<128, 128>(rng);
  std::printf("[%s cp.async] warp-tiled kernel: %lld checks, %lld failures, %lld emulator errors\n",
              eager ? "eager" : "deferred", g_checks - self, g_fail - f0, emu::st().errors - e0);
}

int main() {
  suite(false);
  suite(true);
  std::printf("total: %lld checks, %lld failures, %lld emulator errors; %lld copies, %lld barriers\n", g_checks, g_fail,
              emu::st().errors, emu::st().copies, emu::st().barriers);
  return (g_fail || emu::st().errors) ? 1 : 0;
}
