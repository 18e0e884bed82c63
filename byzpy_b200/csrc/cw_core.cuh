// Device-side core of the coordinate-wise selection kernels, shared by the
// stand-alone launcher (cw_select.cu) and the fused cross-GPU parameter-server
// kernel (fused_ps.cu).
#pragma once
#include "api.h"

namespace bzcw {


constexpr int kThreads = 256;
constexpr float kInf = __builtin_huge_valf();

// PREPAD: slots nt..NP-1 already hold the -inf / +inf padding (the warp-tiled kernel keeps the pad rows
// resident in its shared-memory tile), so the per-slot selects are skipped.
template <int NP, int MODE, bool PREPAD = false>
BZ_HD float cw_pick(float (&v)[NP], const int nt, const int f, const int apad) {
  if constexpr (MODE == BZ_CW_MEAN) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) s += (i < nt) ? v[i] : 0.f;
    return s / (float)nt;
  } else {
    // Pad: `apad` slots of -inf then +inf so that the lower median of the nt
    // real values always lands in the compile-time slot NP/2-1.
    if constexpr (!PREPAD) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (i >= nt) v[i] = (i - nt < apad) ? -kInf : kInf;
      }
    }
    bitonic_sort<NP>(v);
    if constexpr (MODE == BZ_CW_MEDIAN) {
      return v[NP / 2 - 1];
    } else if constexpr (MODE == BZ_CW_TRMEAN) {
      const int lo = apad + f, hi = apad + nt - f;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) s += (i >= lo && i < hi) ? v[i] : 0.f;
      return s / (float)(nt - 2 * f);
    } else {  // BZ_CW_MEAMED
      const float m = v[NP / 2 - 1];
      const int k = nt - f;
      // w[r] = value of real rank r (barrel shift left by apad)
      float w[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) w[i] = v[i];
#pragma unroll
      for (int s = 1; s < NP; s <<= 1) {
        const bool on = (apad & s) != 0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float nxt = (i + s < NP) ? w[i + s] : kInf;
          w[i] = on ? nxt : w[i];
        }
      }
      // u[r] = w[r + k]
      float u[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) u[i] = w[i];
#pragma unroll
      for (int s = 1; s < NP; s <<= 1) {
        const bool on = (k & s) != 0;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float nxt = (i + s < NP) ? u[i + s] : kInf;
          u[i] = on ? nxt : u[i];
        }
      }
      // The k values closest to m form a contiguous window [l, l+k) of the
      // sorted order; slide right while the element entering is closer than
      // the one leaving.
      int l = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const bool c = (i < nt - k) && ((m - w[i]) > (u[i] - m));
        l += c ? 1 : 0;
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) s += (i >= l && i < l + k) ? w[i] : 0.f;
      return s / (float)k;
    }
  }
}

// NC: rows are read-only for the whole kernel (.nc loads); otherwise plain weak loads (see common.cuh)
template <int NP, int V, bool NC = true>
struct VecIO;
template <int NP, bool NC>
struct VecIO<NP, 4, NC> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4][NP], int i, float s) {
    const float4 t = NC ? ldg_stream4(p) : ldg_weak4(p);
    v[0][i] = canon(t.x * s);
    v[1][i] = canon(t.y * s);
    v[2][i] = canon(t.z * s);
    v[3][i] = canon(t.w * s);
  }
};
template <int NP, bool NC>
struct VecIO<NP, 2, NC> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2][NP], int i, float s) {
    const float2 t = NC ? ldg_stream2(p) : ldg_weak2(p);
    v[0][i] = canon(t.x * s);
    v[1][i] = canon(t.y * s);
  }
};
template <int NP, bool NC>
struct VecIO<NP, 1, NC> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1][NP], int i, float s) {
    v[0][i] = canon((NC ? ldg_stream1(p) : ldg_weak1(p)) * s);
  }
};

template <int V>
__device__ __forceinline__ void sgd_apply(const UpdTable& upd, long long idx, const float (&g)[V]) {
  for (int r = 0; r < upd.count; ++r) {
    float* pp = upd.param[r] + idx;
    float* mp = upd.mom[r] ? upd.mom[r] + idx : nullptr;
    float p[V], mo[V];
    if constexpr (V == 4) {
      const float4 t = *reinterpret_cast<const float4*>(pp);
      p[0] = t.x; p[1] = t.y; p[2] = t.z; p[3] = t.w;
      if (mp) {
        const float4 q = *reinterpret_cast<const float4*>(mp);
        mo[0] = q.x; mo[1] = q.y; mo[2] = q.z; mo[3] = q.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) {
        p[c] = pp[c];
        if (mp) mo[c] = mp[c];
      }
    }
#pragma unroll
    for (int c = 0; c < V; ++c) {
      float gg = g[c] + upd.wd * p[c];
      if (mp) {
        mo[c] = upd.mu * mo[c] + gg;
        gg = mo[c];
      }
      p[c] -= upd.lr * gg;
    }
    if constexpr (V == 4) {
      *reinterpret_cast<float4*>(pp) = make_float4(p[0], p[1], p[2], p[3]);
      if (mp) *reinterpret_cast<float4*>(mp) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) {
        pp[c] = p[c];
        if (mp) mp[c] = mo[c];
      }
    }
  }
}


// Values of V coordinates are in v[c][0..n); synthesise virtual rows, run the selection network.
template <int NP, int V, int MODE, bool PREPAD = false>
BZ_HD void cw_finish(float (&v)[V][NP], int n, const VirtRows& virt, int f, float (&res)[V]) {
  const int nv = virt.count;
  const int nt = n + nv;
  const int apad = NP / 2 - 1 - (nt - 1) / 2;
#pragma unroll
  for (int c = 0; c < V; ++c) {
    if (nv > 0) {
      // synthesise the adversary's rows from the honest prefix
      const int nh = virt.n_honest;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) s += (i < nh) ? v[c][i] : 0.f;
      const float mean = s / (float)nh;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float dlt = v[c][i] - mean;
        q += (i < nh) ? dlt * dlt : 0.f;
      }
      const float val = virt.a * mean + virt.b * sqrtf(q / (float)nh);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (i >= n && i < nt) v[c][i] = canon(val);
      }
    }
    res[c] = cw_pick<NP, MODE, PREPAD>(v[c], nt, f, apad);
  }
}

// Load the n real values of V consecutive coordinates straight from the row
// buffers, then finish; returns V results in res[].
template <int NP, int V, int MODE, bool NC = true>
__device__ __forceinline__ void cw_unit(const RowTable& rows, const ScaleTable& scales, int n,
                                        const VirtRows& virt, int f, long long base,
                                        float (&res)[V]) {
  float v[V][NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (i < n) {
      VecIO<NP, V, NC>::load(rows.p[i] + base, v, i, scales.s[i]);
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) v[c][i] = 0.f;
    }
  }
  cw_finish<NP, V, MODE>(v, n, virt, f, res);
}

// ---- cp.async staging: every thread streams ITS OWN next tiles into a private shared-memory
// slot (so no block barrier is needed, only cp.async.wait_group), which keeps two tiles of loads
// in flight per thread while the selection network of the current tile runs, at zero register
// cost.  Slot layout [stage][row][thread][V] is bank-conflict free for both the async writes and
// the read-back.
#ifndef BZ_HOST_EMU      // (the emulator defers the copies to the matching wait, see cuda_host_emu.h)
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#endif

template <int NP, int V>
__device__ __forceinline__ void cw_stage_issue(float* stage, int threads, const RowTable& rows, int n,
                                               long long base) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (i < n) cp_async<V * 4>(stage + ((size_t)i * threads + threadIdx.x) * V, rows.p[i] + base);
  }
}

template <int NP, int V>
__device__ __forceinline__ void cw_stage_read(const float* stage, int threads, const ScaleTable& scales,
                                              int n, float (&v)[V][NP]) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    if (i < n) {
      const float* p = stage + ((size_t)i * threads + threadIdx.x) * V;
      const float sc = scales.s[i];
      if constexpr (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0][i] = canon(t.x * sc);
        v[1][i] = canon(t.y * sc);
        v[2][i] = canon(t.z * sc);
        v[3][i] = canon(t.w * sc);
      } else if constexpr (V == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0][i] = canon(t.x * sc);
        v[1][i] = canon(t.y * sc);
      } else {
        v[0][i] = canon(p[0] * sc);
      }
    } else {
#pragma unroll
      for (int c = 0; c < V; ++c) v[c][i] = 0.f;
    }
  }
}

}  // namespace bzcw
