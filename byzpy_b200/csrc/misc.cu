// Small streaming kernels: attack simulators and the flat optimizer step.
//   colstat   : a*mean + b*std over rows  -> Little / Empire  
//   scale_copy: scale * src               -> SignFlip / Mimic
//   fill      : constant (+inf)           -> InfAttack
//   gaussian  : Philox4x32-10 + Box-Muller N(mu, sigma^2)  -> GaussianAttack (K18)
//   sgd       : fused SGD(+momentum,+wd) over a flat arena for several replicas (K20)
//   u8_affine : uint8 NHWC image batch -> (x - mean[c]) * scale[c] in bf16, one pass (input pipeline)
// Parity: reference attacks/little.py:113-131, empire.py:85-92, sign_flip.py:47-52,
//         inf.py:61-65, gaussian.py:78-83, examples/ps/nodes.py:123-125.
#include <cuda_bf16.h>

#include "api.h"

namespace {

constexpr int kThreads = 256;

inline int grid_for(long long units, int sm_count, int per_sm = 8) {
  long long blocks = (units + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__global__ void __launch_bounds__(kThreads) colstat_kernel(const __grid_constant__ BzColStatArgs a) {
  const int n = a.n;
  const float inv = 1.f / (float)n;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < a.len; j += stride) {
    const long long idx = a.off + j;
    // two passes over the (L1/L2 resident) column keep the variance exact
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += a.rows.p[i][idx] * a.scales.s[i];
    const float mean = s * inv;
    float out = a.a * mean;
    if (a.b != 0.f) {
      float q = 0.f;
      for (int i = 0; i < n; ++i) {
        const float dlt = a.rows.p[i][idx] * a.scales.s[i] - mean;
        q = fmaf(dlt, dlt, q);
      }
      out += a.b * sqrtf(q * inv);
    }
    a.out[idx] = out;
  }
}

__global__ void __launch_bounds__(kThreads) scale_copy_kernel(const float* __restrict__ src,
                                                            float* __restrict__ dst, float scale,
                                                            long long len) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < len; j += stride)
    dst[j] = scale * src[j];
}

__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ dst, float value,
                                                      long long len) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < len; j += stride)
    dst[j] = value;
}

// Philox4x32-10 counter based RNG (Salmon et al.), one 128-bit block -> 4 normals.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t x) {
  // (0, 1]: avoids log(0)
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(kThreads) gaussian_kernel(float* __restrict__ dst, float mu,
                                                          float sigma, unsigned long long seed,
                                                          unsigned long long offset,
                                                          long long len) {
  const long long nblk = (len + 3) / 4;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long b = (long long)blockIdx.x * kThreads + threadIdx.x; b < nblk; b += stride) {
    const unsigned long long ctr = offset + (unsigned long long)b;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    {
      const float r0 = sqrtf(-2.f * logf(u01(c[0])));
      const float r1 = sqrtf(-2.f * logf(u01(c[2])));
      float s0, c0, s1, c1;
      sincospif(2.f * u01(c[1]), &s0, &c0);
      sincospif(2.f * u01(c[3]), &s1, &c1);
      z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long j = b * 4 + k;
      if (j < len) dst[j] = fmaf(sigma, z[k], mu);
    }
  }
}

__global__ void __launch_bounds__(kThreads) sgd_kernel(const float* __restrict__ grad,
                                                     const __grid_constant__ UpdTable upd,
                                                     long long len, int vec) {
  const long long stride = (long long)gridDim.x * kThreads;
  if (vec) {
    const long long nvec = len / 4;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
      const float4 g = *reinterpret_cast<const float4*>(grad + u * 4);
      for (int r = 0; r < upd.count; ++r) {
        float4* pp = reinterpret_cast<float4*>(upd.param[r] + u * 4);
        float4 p = *pp;
        float4 gg = make_float4(g.x + upd.wd * p.x, g.y + upd.wd * p.y, g.z + upd.wd * p.z,
                                g.w + upd.wd * p.w);
        if (upd.mom[r]) {
          float4* mp = reinterpret_cast<float4*>(upd.mom[r] + u * 4);
          float4 m = *mp;
          m.x = upd.mu * m.x + gg.x; m.y = upd.mu * m.y + gg.y;
          m.z = upd.mu * m.z + gg.z; m.w = upd.mu * m.w + gg.w;
          *mp = m;
          gg = m;
        }
        p.x -= upd.lr * gg.x; p.y -= upd.lr * gg.y; p.z -= upd.lr * gg.z; p.w -= upd.lr * gg.w;
        *pp = p;
      }
    }
    // tail
    const long long t0 = nvec * 4;
    const long long j = t0 + (long long)blockIdx.x * kThreads + threadIdx.x;
    if (j < len) {
      const float g = grad[j];
      for (int r = 0; r < upd.count; ++r) {
        const float p = upd.param[r][j];
        float gg = g + upd.wd * p;
        if (upd.mom[r]) {
          const float m = upd.mu * upd.mom[r][j] + gg;
          upd.mom[r][j] = m;
          gg = m;
        }
        upd.param[r][j] = p - upd.lr * gg;
      }
    }
  } else {
    for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < len; j += stride) {
      const float g = grad[j];
      for (int r = 0; r < upd.count; ++r) {
        const float p = upd.param[r][j];
        float gg = g + upd.wd * p;
        if (upd.mom[r]) {
          const float m = upd.mu * upd.mom[r][j] + gg;
          upd.mom[r][j] = m;
          gg = m;
        }
        upd.param[r][j] = p - upd.lr * gg;
      }
    }
  }
}

// 16 pixels-components per thread: one 16-byte load, two 16-byte stores.
struct AffineC {
  float mean[8];
  float scale[8];
};

__global__ void __launch_bounds__(kThreads) u8_affine_kernel(const uint8_t* __restrict__ in,
                                                            __nv_bfloat16* __restrict__ out, long long n,
                                                            int C, const AffineC k) {
  const long long nvec = n >> 4;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
    const uint4 raw = *reinterpret_cast<const uint4*>(in + (u << 4));
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
    int c = (int)((u << 4) % C);
    __nv_bfloat162 o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int b = 2 * q + h;
        const float v = (float)((w[b >> 2] >> (8 * (b & 3))) & 0xffu);
        f[h] = (v - k.mean[c]) * k.scale[c];
        c = (c + 1 == C) ? 0 : c + 1;
      }
      o[q] = __floats2bfloat162_rn(f[0], f[1]);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (u << 4));
    dst[0] = *reinterpret_cast<const uint4*>(&o[0]);
    dst[1] = *reinterpret_cast<const uint4*>(&o[4]);
  }
  // scalar tail
  const long long t0 = nvec << 4;
  for (long long i = t0 + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % C);
    out[i] = __float2bfloat16_rn(((float)in[i] - k.mean[c]) * k.scale[c]);
  }
}

}  // namespace

int bz_colstat(const BzColStatArgs* args, int sm_count, cudaStream_t stream) {
  const BzColStatArgs& a = *args;
  if (a.n < 1 || a.n > BZ_MAXN) return (int)cudaErrorInvalidValue;
  if (a.len <= 0) return 0;
  colstat_kernel<<<grid_for(a.len, sm_count), kThreads, 0, stream>>>(a);
  return (int)cudaGetLastError();
}

int bz_scale_copy(const float* src, float* dst, float scale, long long len, int sm_count,
                  cudaStream_t stream) {
  if (len <= 0) return 0;
  scale_copy_kernel<<<grid_for(len, sm_count), kThreads, 0, stream>>>(src, dst, scale, len);
  return (int)cudaGetLastError();
}

int bz_fill(float* dst, float value, long long len, int sm_count, cudaStream_t stream) {
  if (len <= 0) return 0;
  fill_kernel<<<grid_for(len, sm_count), kThreads, 0, stream>>>(dst, value, len);
  return (int)cudaGetLastError();
}

int bz_gaussian(float* dst, float mu, float sigma, unsigned long long seed,
                unsigned long long offset, long long len, int sm_count, cudaStream_t stream) {
  if (len <= 0) return 0;
  gaussian_kernel<<<grid_for((len + 3) / 4, sm_count), kThreads, 0, stream>>>(dst, mu, sigma, seed,
                                                                             offset, len);
  return (int)cudaGetLastError();
}

int bz_sgd(const float* grad, const UpdTable* upd, long long len, int sm_count,
           cudaStream_t stream) {
  if (len <= 0 || upd->count <= 0) return 0;
  if (upd->count > BZ_MAXR) return (int)cudaErrorInvalidValue;
  bool vec = ((uintptr_t)grad % 16) == 0;
  for (int r = 0; r < upd->count; ++r) {
    vec = vec && ((uintptr_t)upd->param[r] % 16) == 0 &&
          (upd->mom[r] == nullptr || ((uintptr_t)upd->mom[r] % 16) == 0);
  }
  const long long units = vec ? (len + 3) / 4 : len;
  // the vector path's tail needs at least one thread per tail element: grid >= 1 is enough (tail < 4)
  sgd_kernel<<<grid_for(units, sm_count), kThreads, 0, stream>>>(grad, *upd, len, vec ? 1 : 0);
  return (int)cudaGetLastError();
}

int bz_u8_affine(const void* in, void* out, long long n, int C, const float* mean, const float* scale,
                 int sm_count, cudaStream_t stream) {
  if (n < 0 || C < 1 || C > 8) return (int)cudaErrorInvalidValue;
  if (n == 0) return 0;
  if (((uintptr_t)in % 16) != 0 || ((uintptr_t)out % 16) != 0) return (int)cudaErrorMisalignedAddress;
  AffineC k;
  for (int c = 0; c < 8; ++c) {
    k.mean[c] = c < C ? mean[c] : 0.f;
    k.scale[c] = c < C ? scale[c] : 1.f;
  }
  u8_affine_kernel<<<grid_for((n >> 4) + 1, sm_count), kThreads, 0, stream>>>(
      reinterpret_cast<const uint8_t*>(in), reinterpret_cast<__nv_bfloat16*>(out), n, C, k);
  return (int)cudaGetLastError();
}
