// Warp-lockstep host emulation of the coordinate-wise CUDA kernels (no GPU on the build box).
//
// With -DBZ_HOST_EMU the kernel sources under byzpy_b200/csrc compile as plain C++ (g++): this header supplies the
// CUDA vocabulary they use and a small execution model --
//   * every lane of a warp is a coroutine (ucontext) running the kernel body with its own threadIdx; the scheduler
//     advances the 32 lanes phase by phase, a phase ending at __syncwarp(): all lanes must reach the SAME number
//     of barriers (a divergent barrier is reported);
//   * cp.async runs in two modes, and every test runs in both.  DEFERRED: a copy is recorded when issued and
//     performed only when the issuing lane executes the cp.async.wait_group that covers its group -- the latest
//     moment the hardware guarantees, so a read that is not ordered behind the wait (+ barrier, for another lane's
//     copies) sees the poison the shared-memory buffer is filled with (read-after-write hazards).  EAGER: the copy
//     is performed at issue -- the earliest moment the hardware may do it, so a slot that is refilled before all
//     its readers are done is overwritten under them (write-after-read hazards: the scheduler lets a lane run
//     ahead to its next barrier, exactly what an unsynchronised warp may do);
//   * shared memory is one buffer per block, poisoned with NaNs at block start; warps of a block run one after
//     the other (the kernels emulated here have warp- or thread-private shared-memory slots and no block barrier;
//     __syncthreads() aborts).
// What it does not model: memory-system timing, bank conflicts, occupancy -- it is a functional check of index
// maps, pipeline slot rotation, padding, tails and barrier placement.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__
#define __align__(x) __attribute__((aligned(x)))

struct float2 {
  float x, y;
};
struct float4 {
  float x, y, z, w;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };

namespace emu {
struct Idx {
  unsigned x = 0, y = 0, z = 0;
};
struct Copy {
  void* dst;
  const void* src;
  int bytes;
};
struct Lane {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  bool at_barrier = false;
  long long barriers = 0;
  std::vector<std::vector<Copy>> groups;      // committed, not yet completed
  std::vector<Copy> open;                     // issued since the last commit
};
struct State {
  ucontext_t sched;
  std::vector<Lane> lanes;
  int cur = -1;
  std::function<void()> body;
  long long errors = 0;
  long long copies = 0, barriers = 0;
  bool eager = false;                         // cp.async performed at issue instead of at the covering wait
};
inline State& st() {
  static State s;
  return s;
}
inline void fail(const char* what) {
  std::fprintf(stderr, "emulator: %s\n", what);
  ++st().errors;
}
inline void cp_async(void* dst, const void* src, int bytes) {
  if (st().eager) {
    std::memcpy(dst, src, (size_t)bytes);
    ++st().copies;
    st().lanes[st().cur].open.push_back(Copy{nullptr, nullptr, 0});     // still tracked: groups must be waited for
    return;
  }
  st().lanes[st().cur].open.push_back(Copy{dst, src, bytes});
}
inline void commit() {
  Lane& l = st().lanes[st().cur];
  l.groups.push_back(std::move(l.open));
  l.open.clear();
}
inline void wait(int keep) {                  // cp.async.wait_group keep: all but the newest `keep` groups are complete
  Lane& l = st().lanes[st().cur];
  while ((int)l.groups.size() > keep) {
    for (const Copy& c : l.groups.front()) {
      if (c.bytes == 0) continue;
      std::memcpy(c.dst, c.src, (size_t)c.bytes);
      ++st().copies;
    }
    l.groups.erase(l.groups.begin());
  }
}
inline void trampoline() {
  State& s = st();
  s.body();
  Lane& l = s.lanes[s.cur];
  if (!l.open.empty() || !l.groups.empty()) fail("kernel exited with cp.async groups in flight (missing wait_group 0)");
  l.done = true;
  swapcontext(&l.ctx, &s.sched);
}
}  // namespace emu

inline emu::Idx threadIdx, blockIdx, blockDim, gridDim;

inline void __syncwarp() {
  emu::State& s = emu::st();
  emu::Lane& l = s.lanes[s.cur];
  l.at_barrier = true;
  ++l.barriers;
  ++s.barriers;
  swapcontext(&l.ctx, &s.sched);
}
inline void __syncthreads() {
  emu::fail("__syncthreads() is not modelled by this emulator");
  std::abort();
}

namespace emu {
// run `body` (the kernel call) for one warp: lanes warp*32 .. warp*32+31 of the current block
inline void run_warp(int warp, const std::function<void()>& body) {
  State& s = st();
  s.body = body;
  s.lanes.clear();
  s.lanes.resize(32);
  for (int i = 0; i < 32; ++i) {
    Lane& l = s.lanes[i];
    l.stack.resize(1 << 20);
    getcontext(&l.ctx);
    l.ctx.uc_stack.ss_sp = l.stack.data();
    l.ctx.uc_stack.ss_size = l.stack.size();
    l.ctx.uc_link = nullptr;
    makecontext(&l.ctx, (void (*)())trampoline, 0);
  }
  for (;;) {
    int live = 0, finished = 0, waiting = 0;
    for (int i = 0; i < 32; ++i) {
      Lane& l = s.lanes[i];
      if (l.done) continue;
      ++live;
      l.at_barrier = false;
      s.cur = i;
      threadIdx.x = (unsigned)(warp * 32 + i);
      swapcontext(&s.sched, &l.ctx);
      if (l.done) ++finished;
      else if (l.at_barrier) ++waiting;
    }
    if (live == 0) break;
    if (finished != 0 && waiting != 0) {
      fail("divergent __syncwarp(): some lanes exited while others wait at a barrier");
      break;
    }
  }
  s.cur = -1;
}

// <<<grid, threads, smem>>>: blocks and warps one after the other; `smem` is poisoned before every block
inline void launch(unsigned grid, unsigned threads, float* smem, size_t smem_floats, const std::function<void()>& body) {
  gridDim = Idx{grid, 1, 1};
  blockDim = Idx{threads, 1, 1};
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx = Idx{b, 0, 0};
    for (size_t i = 0; i < smem_floats; ++i) smem[i] = std::numeric_limits<float>::quiet_NaN();
    for (unsigned w = 0; w < threads / 32; ++w) run_warp((int)w, body);
  }
}
}  // namespace emu

// ---- the memory-access helpers of common.cuh / cw_core.cuh (PTX there, plain accesses here)
inline float4 ldg_stream4(const float* p) { float4 r; std::memcpy(&r, p, 16); return r; }
inline float2 ldg_stream2(const float* p) { float2 r; std::memcpy(&r, p, 8); return r; }
inline float ldg_stream1(const float* p) { return *p; }
inline float4 ldg_weak4(const float* p) { return ldg_stream4(p); }
inline float2 ldg_weak2(const float* p) { return ldg_stream2(p); }
inline float ldg_weak1(const float* p) { return *p; }
inline float4 ldg_cg4(const float* p) { return ldg_stream4(p); }
inline float ldg_cg1(const float* p) { return *p; }
inline void stg_stream4(float* p, float4 v) { std::memcpy(p, &v, 16); }
inline void stg_stream1(float* p, float v) { *p = v; }
inline void stg_multimem4(float* p, float4 v) { std::memcpy(p, &v, 16); }
inline void stg_multimem1(float* p, float v) { *p = v; }

template <int BYTES>
inline void cp_async(void* smem, const void* gmem) {
  if (((uintptr_t)smem % BYTES) != 0 || ((uintptr_t)gmem % BYTES) != 0) emu::fail("misaligned cp.async");
  emu::cp_async(smem, gmem, BYTES);
}
inline void cp_async_commit() { emu::commit(); }
template <int N>
inline void cp_async_wait() {
  emu::wait(N);
}
