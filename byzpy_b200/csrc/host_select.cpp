// Host (CPU) coordinate-wise selection: the same idea as cw_select.cu, on SIMD lanes instead of CUDA
// threads.  A tile of TILE consecutive coordinates of all n rows is copied into an L1-resident buffer
// (NaN canonicalised to +inf, optional per-row scale applied), sorted along the row axis by a
// data-independent merge-exchange network whose comparators are plain min/max over TILE-wide rows (the
// compiler turns them into packed min/max), and the statistic is read off the sorted tile.  Tiles are
// independent, so they are split over threads.
//
// This replaces torch.stack + torch.sort / median / kthvalue on the CPU path of CoordinateWiseMedian,
// CoordinateWiseTrimmedMean and MeanOfMedians (reference aggregators/coordinate_wise/median.py:102-106,
// trimmed_mean.py:110-115, mean_of_medians.py:71-81): no (n, d) stack is materialised and the per-column
// strided sort disappears.
#include "host_select.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr int TILE = 64;       // coordinates per tile: n * TILE * 4 B = 16 KB at n = 64
constexpr int MAX_ROWS = 1024;  // network cache bound (the operators use n <= a few hundred)

struct Network {
  std::vector<uint16_t> lo, hi;  // comparator k orders rows (lo[k], hi[k]) ascending
};

// Knuth's merge exchange (TAOCP 5.2.2, Algorithm M): a sorting network for any n, O(n log^2 n) comparators.
Network build_network(int n) {
  Network net;
  if (n < 2) return net;
  int t = 0;
  while ((1 << t) < n) ++t;
  for (int p = 1 << (t - 1); p >= 1; p >>= 1) {
    int q = 1 << (t - 1), r = 0, d = p;
    while (true) {
      for (int i = 0; i + d < n; ++i)
        if ((i & p) == r) {
          net.lo.push_back((uint16_t)i);
          net.hi.push_back((uint16_t)(i + d));
        }
      if (q == p) break;
      d = q - p;
      q >>= 1;
      r = p;
    }
  }
  return net;
}

const Network& network_for(int n) {
  static std::mutex mu;
  static std::vector<Network> cache(MAX_ROWS + 1);
  static std::vector<char> have(MAX_ROWS + 1, 0);
  std::lock_guard<std::mutex> lock(mu);
  if (!have[n]) {
    cache[n] = build_network(n);
    have[n] = 1;
  }
  return cache[n];
}

#if defined(__GNUC__) && defined(__x86_64__)
#define BZ_HOST_SIMD __attribute__((target("avx2")))
#else
#define BZ_HOST_SIMD
#endif

// --- tile kernels, compiled twice: baseline ISA and AVX2 (picked once at run time) -----------------
#define BZ_DEFINE_TILE_KERNELS(SUFFIX, ATTR)                                                        \
  ATTR void load_tile_##SUFFIX(float* buf, const float* const* rows, const float* scales, int n,   \
                               int64_t start, int width) {                                         \
    const float inf = std::numeric_limits<float>::infinity();                                      \
    for (int i = 0; i < n; ++i) {                                                                  \
      const float* src = rows[i] + start;                                                          \
      float* dst = buf + (int64_t)i * TILE;                                                        \
      const float s = scales ? scales[i] : 1.0f;                                                   \
      if (scales) {                                                                                \
        for (int j = 0; j < width; ++j) {                                                          \
          float v = src[j] * s;                                                                    \
          dst[j] = (v != v) ? inf : v;                                                             \
        }                                                                                          \
      } else {                                                                                     \
        for (int j = 0; j < width; ++j) {                                                          \
          float v = src[j];                                                                        \
          dst[j] = (v != v) ? inf : v;                                                             \
        }                                                                                          \
      }                                                                                            \
      for (int j = width; j < TILE; ++j) dst[j] = 0.0f;                                            \
    }                                                                                              \
  }                                                                                                \
  ATTR void sort_tile_##SUFFIX(float* buf, const uint16_t* lo, const uint16_t* hi, size_t count) { \
    for (size_t k = 0; k < count; ++k) {                                                           \
      float* __restrict a = buf + (int64_t)lo[k] * TILE;                                           \
      float* __restrict b = buf + (int64_t)hi[k] * TILE;                                           \
      for (int j = 0; j < TILE; ++j) {                                                             \
        const float x = a[j], y = b[j];                                                            \
        a[j] = x < y ? x : y;                                                                      \
        b[j] = x < y ? y : x;                                                                      \
      }                                                                                            \
    }                                                                                              \
  }

BZ_DEFINE_TILE_KERNELS(base, )
BZ_DEFINE_TILE_KERNELS(avx2, BZ_HOST_SIMD)

bool use_avx2() {
#if defined(__GNUC__) && defined(__x86_64__)
  static const bool ok = __builtin_cpu_supports("avx2");
  return ok;
#else
  return false;
#endif
}

// statistic of one sorted tile (rows ascending along the row axis) -> out[start .. start + width)
void emit_tile(const float* buf, int n, int mode, int f, float* out, int64_t start, int width) {
  const int mid = (n - 1) / 2;
  if (mode == BZ_HOST_MEDIAN) {
    const float* m = buf + (int64_t)mid * TILE;
    for (int j = 0; j < width; ++j) out[start + j] = m[j];
    return;
  }
  if (mode == BZ_HOST_TRMEAN) {
    double acc[TILE];
    for (int j = 0; j < TILE; ++j) acc[j] = 0.0;
    for (int i = f; i < n - f; ++i) {
      const float* r = buf + (int64_t)i * TILE;
      for (int j = 0; j < TILE; ++j) acc[j] += (double)r[j];
    }
    const double inv = 1.0 / (double)(n - 2 * f);
    for (int j = 0; j < width; ++j) out[start + j] = (float)(acc[j] * inv);
    return;
  }
  // mean of medians: the k = n - f values closest to the median are a contiguous window of the sorted
  // column; its left edge is the number of positions i for which dropping S[i] beats dropping S[i + k]
  const int k = n - f;
  const float* m = buf + (int64_t)mid * TILE;
  for (int j = 0; j < width; ++j) {
    int left = 0;
    const float mj = m[j];
    for (int i = 0; i < n - k; ++i) {
      const float a = mj - buf[(int64_t)i * TILE + j];
      const float b = buf[(int64_t)(i + k) * TILE + j] - mj;
      left += (a > b) ? 1 : 0;
    }
    double acc = 0.0;
    for (int i = left; i < left + k; ++i) acc += (double)buf[(int64_t)i * TILE + j];
    out[start + j] = (float)(acc / (double)k);
  }
}

void run_range(const float* const* rows, const float* scales, int n, int mode, int f, float* out,
               int64_t tile_begin, int64_t tile_end, int64_t d, const Network& net) {
  std::vector<float> storage((size_t)n * TILE + 16);
  float* buf = storage.data();
  // 64-byte alignment keeps every tile row on its own cache lines
  buf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(buf) + 63) & ~uintptr_t(63));
  const bool avx2 = use_avx2();
  for (int64_t t = tile_begin; t < tile_end; ++t) {
    const int64_t start = t * TILE;
    const int width = (int)std::min<int64_t>(TILE, d - start);
    if (avx2) {
      load_tile_avx2(buf, rows, scales, n, start, width);
      sort_tile_avx2(buf, net.lo.data(), net.hi.data(), net.lo.size());
    } else {
      load_tile_base(buf, rows, scales, n, start, width);
      sort_tile_base(buf, net.lo.data(), net.hi.data(), net.lo.size());
    }
    emit_tile(buf, n, mode, f, out, start, width);
  }
}

void mean_range(const float* const* rows, const float* scales, int n, float* out, int64_t begin,
                int64_t end) {
  const float inf = std::numeric_limits<float>::infinity();
  constexpr int B = 256;
  double acc[B];
  for (int64_t s = begin; s < end; s += B) {
    const int w = (int)std::min<int64_t>(B, end - s);
    for (int j = 0; j < w; ++j) acc[j] = 0.0;
    for (int i = 0; i < n; ++i) {
      const float* r = rows[i] + s;
      const float sc = scales ? scales[i] : 1.0f;
      for (int j = 0; j < w; ++j) {
        float v = r[j] * sc;
        acc[j] += (double)((v != v) ? inf : v);
      }
    }
    for (int j = 0; j < w; ++j) out[s + j] = (float)(acc[j] / (double)n);
  }
}

// a * mean + b * (population std) per coordinate, two in-cache sweeps over a block of columns (Little /
// Empire attacks: reference attacks/little.py:113-131, empire.py:85-92).  No NaN canonicalisation: the
// statistic of a column holding NaN / inf is NaN / inf as in the PyTorch formula.
void colstat_range(const float* const* rows, const float* scales, int n, double a, double b, float* out,
                   int64_t begin, int64_t end) {
  constexpr int B = 256;
  double mean[B], var[B];
  for (int64_t s = begin; s < end; s += B) {
    const int w = (int)std::min<int64_t>(B, end - s);
    for (int j = 0; j < w; ++j) mean[j] = var[j] = 0.0;
    for (int i = 0; i < n; ++i) {
      const float* r = rows[i] + s;
      const float sc = scales ? scales[i] : 1.0f;
      for (int j = 0; j < w; ++j) mean[j] += (double)(r[j] * sc);
    }
    for (int j = 0; j < w; ++j) mean[j] /= (double)n;
    if (b != 0.0) {
      for (int i = 0; i < n; ++i) {
        const float* r = rows[i] + s;
        const float sc = scales ? scales[i] : 1.0f;
        for (int j = 0; j < w; ++j) {
          const double c = (double)(r[j] * sc) - mean[j];
          var[j] += c * c;
        }
      }
      for (int j = 0; j < w; ++j) out[s + j] = (float)(a * mean[j] + b * std::sqrt(var[j] / (double)n));
    } else {
      for (int j = 0; j < w; ++j) out[s + j] = (float)(a * mean[j]);
    }
  }
}

template <typename Fn>
void split_columns(int64_t d, int threads, Fn&& fn) {
  const int64_t blocks = (d + 4095) / 4096;
  const int use = (int)std::max<int64_t>(1, std::min<int64_t>(threads, blocks / 4));
  if (use <= 1) {
    fn(0, d);
    return;
  }
  std::vector<std::thread> pool;
  for (int w = 0; w < use; ++w) {
    const int64_t b = blocks * w / use * 4096, e = std::min<int64_t>(d, blocks * (w + 1) / use * 4096);
    pool.emplace_back(fn, b, e);
  }
  for (auto& th : pool) th.join();
}

}  // namespace

int bz_host_colstat(const float* const* rows, const float* scales, int n, int64_t d, double a, double b,
                    float* out, int threads) {
  if (n < 1 || d < 0) return 1;
  if (d == 0) return 0;
  split_columns(d, std::max(1, threads), [=](int64_t lo, int64_t hi) {
    colstat_range(rows, scales, n, a, b, out, lo, hi);
  });
  return 0;
}

int bz_host_network_size(int n) {
  if (n < 1 || n > MAX_ROWS) return -1;
  return (int)network_for(n).lo.size();
}

int bz_host_cw_select(const float* const* rows, const float* scales, int n, int64_t d, int mode, int f,
                      float* out, int threads) {
  if (n < 1 || n > MAX_ROWS || d < 0) return 1;
  if (mode == BZ_HOST_TRMEAN && !(f >= 0 && 2 * f < n)) return 2;
  if (mode == BZ_HOST_MEAMED && !(f >= 0 && f < n)) return 2;
  if (mode < BZ_HOST_MEDIAN || mode > BZ_HOST_MEAN) return 3;
  if (d == 0) return 0;
  threads = std::max(1, threads);
  if (mode == BZ_HOST_MEAN) {
    const int64_t blocks = (d + 4095) / 4096;
    const int use = (int)std::min<int64_t>(threads, blocks);
    if (use <= 1) {
      mean_range(rows, scales, n, out, 0, d);
      return 0;
    }
    std::vector<std::thread> pool;
    for (int w = 0; w < use; ++w) {
      const int64_t b = blocks * w / use * 4096, e = std::min<int64_t>(d, blocks * (w + 1) / use * 4096);
      pool.emplace_back(mean_range, rows, scales, n, out, b, e);
    }
    for (auto& th : pool) th.join();
    return 0;
  }
  const Network& net = network_for(n);
  const int64_t tiles = (d + TILE - 1) / TILE;
  // below ~64 tiles per thread the spawn cost is visible; small inputs run on the caller's thread
  const int use = (int)std::max<int64_t>(1, std::min<int64_t>(threads, tiles / 64));
  if (use <= 1) {
    run_range(rows, scales, n, mode, f, out, 0, tiles, d, net);
    return 0;
  }
  std::vector<std::thread> pool;
  for (int w = 0; w < use; ++w) {
    const int64_t b = tiles * w / use, e = tiles * (w + 1) / use;
    pool.emplace_back(run_range, rows, scales, n, mode, f, out, b, e, d, std::cref(net));
  }
  for (auto& th : pool) th.join();
  return 0;
}
