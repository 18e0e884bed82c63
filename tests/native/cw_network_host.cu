// Host build of the register selection code of the coordinate-wise kernels (csrc/cw_core.cuh: the odd-even
// merge network, the -inf / +inf padding that pins the lower median to a compile-time slot, trimmed mean,
// mean-of-medians, the synthesised Little / Empire rows, NaN canonicalisation) checked against a plain sort.
// Built and run by tests/test_cw_network_host.py with nvcc on a box WITHOUT a GPU: the functions are
// __host__ __device__, so this executes exactly the source the kernels compile.
//
//   nvcc -std=c++17 -O1 -I byzpy_b200/csrc tests/native/cw_network_host.cu -o /tmp/cw_network_host && /tmp/cw_network_host
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "cw_core.cuh"

using namespace bzcw;

static int failures = 0;
static long long checks = 0;

static float ref_value(std::vector<float> x, int mode, int f) {      // x: the nt values, NaN already +inf
  const int nt = (int)x.size();
  std::sort(x.begin(), x.end());
  if (mode == BZ_CW_MEDIAN) return x[(nt - 1) / 2];
  if (mode == BZ_CW_MEAN) {
    float s = 0.f;
    for (float v : x) s += v;
    return s / (float)nt;
  }
  if (mode == BZ_CW_TRMEAN) {
    float s = 0.f;
    for (int i = f; i < nt - f; ++i) s += x[i];
    return s / (float)(nt - 2 * f);
  }
  // mean of the k = nt - f values closest to the lower median (ties: the sliding window of the kernel keeps
  // the left-most window among equals, which is what this scan does too)
  const float m = x[(nt - 1) / 2];
  const int k = nt - f;
  int l = 0;
  while (l + k < nt && (m - x[l]) > (x[l + k] - m)) ++l;
  float s = 0.f;
  for (int i = l; i < l + k; ++i) s += x[i];
  return s / (float)k;
}

static bool same(float a, float b) {
  if (std::isnan(a) || std::isnan(b)) return std::isnan(a) && std::isnan(b);
  if (std::isinf(a) || std::isinf(b)) return a == b;
  return std::fabs(a - b) <= 1e-5f * (1.f + std::fabs(b));
}

template <int NP, int MODE, bool PREPAD>
static void check_case(std::mt19937& rng, int n, int nv, int f, int flavour) {
  const int nt = n + nv;
  std::normal_distribution<float> nd(0.f, 1.f);
  float v[1][NP];
  std::vector<float> real(n);
  for (int i = 0; i < n; ++i) {
    float x = nd(rng);
    if (flavour == 1 && (rng() % 4) == 0) x = (float)((int)(x * 2.f));          // many ties
    if (flavour == 2 && (rng() % 7) == 0) x = std::nanf("");                    // NaN -> +inf
    if (flavour == 3 && (rng() % 5) == 0) x = (rng() & 1) ? INFINITY : -INFINITY;
    real[i] = x;
  }
  VirtRows virt{nv, std::max(1, n - 1), 0.7f, -1.3f};
  if (nv == 0) virt = VirtRows{0, 0, 0.f, 0.f};
  for (int i = 0; i < NP; ++i) v[0][i] = 0.f;
  for (int i = 0; i < n; ++i) v[0][i] = (flavour == 2 && PREPAD) ? canon_min(real[i]) : canon(real[i]);
  if (PREPAD) {   // what the warp-tiled kernel keeps resident in its shared-memory tile
    const int apad = NP / 2 - 1 - (nt - 1) / 2;
    for (int i = nt; i < NP; ++i) v[0][i] = (i - nt < apad) ? -INFINITY : INFINITY;
  } else {
    for (int i = n; i < NP; ++i) v[0][i] = 12345.f;    // garbage the padding must overwrite
  }
  // oracle
  std::vector<float> x(nt);
  for (int i = 0; i < n; ++i) x[i] = std::isnan(real[i]) ? INFINITY : real[i];
  if (nv > 0) {
    const int nh = virt.n_honest;
    float s = 0.f;
    for (int i = 0; i < nh; ++i) s += x[i];
    const float mean = s / (float)nh;
    float q = 0.f;
    for (int i = 0; i < nh; ++i) q += (x[i] - mean) * (x[i] - mean);
    float val = virt.a * mean + virt.b * std::sqrt(q / (float)nh);
    if (std::isnan(val)) val = INFINITY;
    for (int i = n; i < nt; ++i) x[i] = val;
  }
  const float want = ref_value(x, MODE, f);
  float res[1];
  cw_finish<NP, 1, MODE, PREPAD>(v, n, virt, f, res);
  ++checks;
  if (!same(res[0], want)) {
    if (failures < 20)
      std::printf("MISMATCH NP=%d mode=%d prepad=%d n=%d nv=%d f=%d flavour=%d: got %.9g want %.9g\n", NP, MODE,
                  (int)PREPAD, n, nv, f, flavour, res[0], want);
    ++failures;
  }
}

template <int NP, int MODE, bool PREPAD>
static void sweep(std::mt19937& rng, int reps) {
  const int lo = NP / 2 + 1 > 1 ? NP / 2 + 1 : 1;
  for (int nt = (NP == 2 ? 1 : lo); nt <= NP; ++nt) {
    for (int nv = 0; nv <= 2 && nv < nt; ++nv) {
      const int n = nt - nv;
      if (nv > 0 && n < 2) continue;
      std::vector<int> fs;
      if (MODE == BZ_CW_TRMEAN) fs = {0, 1, (nt - 1) / 2 > 0 ? (nt - 1) / 2 - 0 : 0, (nt - 1) / 4};
      else if (MODE == BZ_CW_MEAMED) fs = {0, 1, nt / 3, nt - 1};
      else fs = {0};
      for (int f : fs) {
        if (MODE == BZ_CW_TRMEAN && !(f >= 0 && 2 * f < nt)) continue;
        if (MODE == BZ_CW_MEAMED && !(f >= 0 && f < nt)) continue;
        for (int flavour = 0; flavour < 4; ++flavour) {
          if (nv > 0 && flavour >= 2) continue;        // non-finite honest rows make the synthesised value NaN / inf
          for (int r = 0; r < reps; ++r) check_case<NP, MODE, PREPAD>(rng, n, nv, f, flavour);
        }
      }
    }
  }
}

template <int NP>
static void all_modes(std::mt19937& rng, int reps) {
  sweep<NP, BZ_CW_MEDIAN, false>(rng, reps);
  sweep<NP, BZ_CW_TRMEAN, false>(rng, reps);
  sweep<NP, BZ_CW_MEAMED, false>(rng, reps);
  sweep<NP, BZ_CW_MEDIAN, true>(rng, reps);
  sweep<NP, BZ_CW_TRMEAN, true>(rng, reps);
  sweep<NP, BZ_CW_MEAMED, true>(rng, reps);
}

// 0-1 principle for the full sort: every 0/1 input of NP <= 16 bits, random 0/1 inputs above
template <int NP>
static void zero_one(std::mt19937& rng) {
  const long long total = NP <= 16 ? (1ll << NP) : 200000;
  for (long long t = 0; t < total; ++t) {
    float v[NP];
    int ones = 0;
    for (int i = 0; i < NP; ++i) {
      const int bit = NP <= 16 ? (int)((t >> i) & 1) : (int)(rng() & 1);
      v[i] = (float)bit;
      ones += bit;
    }
    bitonic_sort<NP>(v);
    ++checks;
    for (int i = 0; i < NP; ++i) {
      if (v[i] != (i >= NP - ones ? 1.f : 0.f)) {
        if (failures < 20) std::printf("0-1 principle violated NP=%d input %lld\n", NP, t);
        ++failures;
        break;
      }
    }
  }
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? std::atoi(argv[1]) : 3;
  std::mt19937 rng(1234);
  zero_one<2>(rng); zero_one<4>(rng); zero_one<8>(rng); zero_one<16>(rng);
  zero_one<32>(rng); zero_one<64>(rng); zero_one<128>(rng);
  all_modes<2>(rng, reps * 8); all_modes<4>(rng, reps * 8); all_modes<8>(rng, reps * 4); all_modes<16>(rng, reps * 2);
  all_modes<32>(rng, reps); all_modes<64>(rng, reps); all_modes<128>(rng, reps);
  std::printf("%lld checks, %d failures\n", checks, failures);
  return failures ? 1 : 0;
}
