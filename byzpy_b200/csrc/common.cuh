// Shared device helpers for the byzpy_b200 sm_100a kernels.
//
// Conventions used by every kernel in this directory:
//   * a "row" is one flattened gradient / parameter vector of length d (fp32);
//   * the n rows of the logical (n, d) matrix are NOT required to be contiguous:
//     kernels take a RowTable of n base pointers, so rows may live in different
//     allocations, in a flat arena, or in a *peer GPU's* HBM mapped over NVLink;
//   * per-row scale factors (sign-flip, clipping, ...) are folded into the load.
#pragma once
#ifdef BZ_HOST_EMU
// warp-lockstep host emulation of the kernels (tests/native/cuda_host_emu.h, tests/test_cw_kernels_emulated.py):
// the CUDA runtime types, thread indices, memory-access helpers and cp.async come from the emulator
#include "cuda_host_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define BZ_MAXN 128   // max rows handled by the register / smem resident kernels
#define BZ_MAXR 16    // max model replicas updated by a fused optimizer epilogue

struct RowTable {
  const float* p[BZ_MAXN];
};
struct ScaleTable {
  float s[BZ_MAXN];
};

// Optional fused SGD(+momentum, +weight-decay) epilogue applied to `count`
// local model replicas with the freshly aggregated gradient.
struct UpdTable {
  float* param[BZ_MAXR];
  float* mom[BZ_MAXR];
  int count;
  float lr, mu, wd;
};

// Virtual (synthesised) rows: `count` identical rows whose value per coordinate
// is a*mean(h) + b*std(h) over the first `n_honest` real rows.  Little ("a
// little is enough") and Empire attacks are exactly this, so the
// omniscient adversary never materialises its vector.
struct VirtRows {
  int count;
  int n_honest;
  float a, b;
};

#ifndef BZ_HOST_EMU
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
               : "=f"(r.x), "=f"(r.y)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// Streaming loads WITHOUT the non-coherent qualifier: for rows that a peer GPU may still have been
// writing after this kernel became resident (the fused rounds wait for the producer's flag with
// ld.acquire.sys inside the kernel; PTX only allows .nc on memory that is read-only for the whole
// kernel).  Same LDG path and L1 policy as the .nc variants.
__device__ __forceinline__ float4 ldg_weak4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ float2 ldg_weak2(const float* p) {
  float2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ldg_weak1(const float* p) {
  float r;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
// Coherent (L2) loads for data produced by a peer GPU *during* this kernel.
__device__ __forceinline__ float4 ldg_cg4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_cg1(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// NVLS multicast store: one store to the multicast alias lands in every bound GPU's HBM (the
// NVSwitch replicates it), replacing `world` peer stores.
__device__ __forceinline__ void stg_multimem4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_multimem1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

#endif  // !BZ_HOST_EMU

// The register-resident selection code below is plain arithmetic, so it is compiled for the host as
// well: tests/test_cw_network_host.py builds it with nvcc and checks every network size, padding and
// mode against a sort on the CPU (the box that builds the extension has no GPU).
#define BZ_HD __host__ __device__ __forceinline__

// NaN is canonicalised to +inf on load: a NaN coordinate is treated as an
// extreme outlier (sorts last, like torch.sort) instead of poisoning min/max.
BZ_HD float canon(float x) { return (x != x) ? __builtin_huge_valf() : x; }
// The same map in ONE alu instruction: IEEE minNum returns the numeric operand, so min(NaN, +inf) = +inf
// and min(x, +inf) = x for every other x (FMNMX instead of FSETP + FSEL).
BZ_HD float canon_min(float x) { return fminf(x, __builtin_huge_valf()); }

// Fully unrolled bitonic sorting network over a register array (ascending).
// All indices are compile-time after unrolling, so v[] stays in registers, and
// the compiler dead-code-eliminates compare-exchanges whose outputs are unused
// (e.g. when only the median slot is read).
template <int NP>
BZ_HD void bitonic_sort(float (&v)[NP]) {
  // Batcher's odd-even merge sort (19 / 63 / 191 / 543 / 1471 compare-exchanges for
  // NP = 8 / 16 / 32 / 64 / 128, ~20 % fewer than the bitonic network the name recalls).
  // Every comparator sorts ascending, so each is exactly one FMNMX pair on the alu pipe.
#pragma unroll
  for (int p = 1; p < NP; p <<= 1) {
#pragma unroll
    for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
      for (int x = 0; x < NP; ++x) {
        const int r = k % p;
        if (x >= r && ((x - r) % (2 * k)) < k && x + k < NP && (x / (2 * p)) == ((x + k) / (2 * p))) {
          const float a = v[x], b = v[x + k];
          v[x] = fminf(a, b);
          v[x + k] = fmaxf(a, b);
        }
      }
    }
  }
}

#ifndef BZ_HOST_EMU
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif

#define BZ_CUDA_CHECK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return (int)_e;                    \
  } while (0)
