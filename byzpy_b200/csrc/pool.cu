// 3x3 / stride 2 / pad 1 max pooling for NHWC bf16 activations (the ResNet stem pool).
//
// ATen's max_pool_{forward,backward}_nhwc take 107 us / 227 us on the 32x64x112x112 stem activation
// (profiles/bench_log.md) -- the backward alone is 6 % of a replica's step.  Both directions here
// are pure streaming kernels with 16-byte accesses (8 channels per thread):
//   forward   y = max over the window, plus a one-byte window position (kh*3+kw) per output element;
//             scan order and the strict '>' (NaN propagates) match ATen, so ties pick the same element
//   backward  GATHER form: each 2x2 block of input pixels loads the <= 4 windows that cover it once and
//             routes every dy to the pixel whose recorded position matches -- no atomics, deterministic.
#include <cuda_bf16.h>
#include <stdint.h>

#include "pool.h"

namespace {

constexpr int kThreads = 256;

// Eight bf16 values moved as ONE 128-bit access.  (A struct of four __nv_bfloat162 is copied member by
// member -- four LDG.32 -- which capped the first version of these kernels at a quarter of the load
// width; ncu: 1.75 TB/s on the 51 MB stem activation.)
struct alignas(16) Bf8 {
  uint4 u;
};

struct alignas(8) Idx8 {
  uint8_t b[8];
};

__device__ __forceinline__ void unpack(const Bf8& p, float (&f)[8]) {
  const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // bf16 -> fp32 is a 16-bit shift
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ Bf8 pack(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  Bf8 p;
  p.u = make_uint4(w[0], w[1], w[2], w[3]);
  return p;
}
__device__ __forceinline__ Bf8 ld8(const __nv_bfloat16* p) {
  Bf8 r;
  r.u = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const Bf8& v) { *reinterpret_cast<uint4*>(p) = v.u; }

__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                              __nv_bfloat16* __restrict__ y,
                                                              uint8_t* __restrict__ idx, int N, int H, int W,
                                                              int C, int Ho, int Wo) {
  const int cg = C >> 3;
  const long long total = (long long)N * Ho * Wo * cg;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < total; u += stride) {
    const int g = (int)(u % cg);
    long long t = u / cg;
    const int ow = (int)(t % Wo);
    t /= Wo;
    const int oh = (int)(t % Ho);
    const int n = (int)(t / Ho);
    // all (<= 9) window loads are issued before the first compare
    Bf8 win[9];
    bool ok[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ih = 2 * oh - 1 + kh, iw = 2 * ow - 1 + kw;
        const bool in = (ih >= 0 && ih < H && iw >= 0 && iw < W);
        ok[kh * 3 + kw] = in;
        win[kh * 3 + kw] = ld8(x + (((long long)n * H + (in ? ih : 0)) * W + (in ? iw : 0)) * C + g * 8);
      }
    }
    float m[8];
    uint8_t am[8];
    bool first = true;
#pragma unroll
    for (int code = 0; code < 9; ++code) {
      if (!ok[code]) continue;
      float f[8];
      unpack(win[code], f);
      if (first) {
        // ATen starts from -inf with the first in-range position as the index
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          m[k] = -__builtin_huge_valf();
          am[k] = (uint8_t)code;
        }
        first = false;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (f[k] > m[k] || f[k] != f[k]) {
          m[k] = f[k];
          am[k] = (uint8_t)code;
        }
      }
    }
    st8(y + u * 8, pack(m));
    Idx8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.b[k] = am[k];
    *reinterpret_cast<Idx8*>(idx + u * 8) = o;
  }
}

// One thread per 2x2 block of input pixels (x 8 channels): the block is covered by the <= 4 windows
// (a .. a+1) x (b .. b+1), whose (index, dy) pairs are loaded ONCE and routed to the four pixels
// (the first gather version re-read them per pixel: 4x the load traffic, 54 us on the stem).
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                              const uint8_t* __restrict__ idx,
                                                              __nv_bfloat16* __restrict__ dx, int N, int H,
                                                              int W, int C, int Ho, int Wo) {
  const int cg = C >> 3;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
  const long long total = (long long)N * Hb * Wb * cg;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < total; u += stride) {
    const int g = (int)(u % cg);
    long long t = u / cg;
    const int b = (int)(t % Wb);
    t /= Wb;
    const int a = (int)(t % Hb);
    const int n = (int)(t / Hb);
    float acc[2][2][8];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[p][q][k] = 0.f;
    // the (<= 4) windows' index + dy loads go out first, then they are routed
    Idx8 wid[2][2];
    Bf8 wdy[2][2];
    bool wok[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int oh = a + i, ow = b + j;
        const bool in = (oh < Ho && ow < Wo);
        wok[i][j] = in;
        const long long o = (((long long)n * Ho + (in ? oh : 0)) * Wo + (in ? ow : 0)) * cg + g;
        wid[i][j] = *reinterpret_cast<const Idx8*>(idx + o * 8);
        wdy[i][j] = ld8(dy + o * 8);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!wok[i][j]) continue;
        const Idx8 id = wid[i][j];
        float d[8];
        unpack(wdy[i][j], d);
        // pixel (p, q) of the block lies in window (oh, ow) iff (i == 0 || p == 1) && (j == 0 || q == 1);
        // its position inside the window is kh = p + 1 - 2i, kw = q + 1 - 2j
#pragma unroll
        for (int p = i; p < 2; ++p) {
#pragma unroll
          for (int q = j; q < 2; ++q) {
            const uint8_t code = (uint8_t)((p + 1 - 2 * i) * 3 + (q + 1 - 2 * j));
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[p][q][k] += (id.b[k] == code) ? d[k] : 0.f;
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int ih = 2 * a + p;
      if (ih >= H) continue;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int iw = 2 * b + q;
        if (iw >= W) continue;
        st8(dx + ((((long long)n * H + ih) * W + iw) * cg + g) * 8, pack(acc[p][q]));
      }
    }
  }
}

int blocks_for(long long total, int sm_count) {
  long long b = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int bz_maxpool3x3s2_forward(const void* x, void* y, void* idx, int N, int H, int W, int C, int sm_count,
                            cudaStream_t stream) {
  if (N < 1 || H < 1 || W < 1 || C < 8 || (C % 8) != 0) return (int)cudaErrorInvalidValue;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;  // floor((H + 2 - 3) / 2) + 1
  const long long total = (long long)N * Ho * Wo * (C >> 3);
  maxpool_fwd_kernel<<<blocks_for(total, sm_count), kThreads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<uint8_t*>(idx), N, H, W, C, Ho, Wo);
  return (int)cudaGetLastError();
}

int bz_maxpool3x3s2_backward(const void* dy, const void* idx, void* dx, int N, int H, int W, int C,
                             int sm_count, cudaStream_t stream) {
  if (N < 1 || H < 1 || W < 1 || C < 8 || (C % 8) != 0) return (int)cudaErrorInvalidValue;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C >> 3);
  maxpool_bwd_kernel<<<blocks_for(total, sm_count), kThreads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const uint8_t*>(idx),
      reinterpret_cast<__nv_bfloat16*>(dx), N, H, W, C, Ho, Wo);
  return (int)cudaGetLastError();
}
