// Native runtime of byzpy_b200 (host side, C++17).
//
// Symmetric memory: every rank cudaMalloc's identically sized arenas, exports
// them as CUDA IPC handles, and maps every peer's arena into its own address
// space.  The kernels in fused_ps.cu then address peer HBM directly
// (ld.global / st.global over NVLink 5 through NVSwitch); NCCL/Gloo are used
// only to exchange the 64-byte handles at start-up.  This replaces the
// reference's POSIX-shm store (reference engine/storage/shared_store.py:21-54)
// and its pickle/UCX tensor transport (engine/actor/transports/ucx.py:225-270).
#include "runtime.h"

#include <pybind11/stl.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "fused_ps.h"

namespace py = pybind11;

namespace {

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
template <typename T>
T* as_ptr(uint64_t v) {
  return reinterpret_cast<T*>(static_cast<uintptr_t>(v));
}
cudaStream_t as_stream(uint64_t s) { return reinterpret_cast<cudaStream_t>(static_cast<uintptr_t>(s)); }

// The part of the fused-round argument block shared by the coordinate-wise and weighted-sum kernels.
void fill_common(BzFusedPsArgs& a, const std::vector<uint64_t>& rows, const std::vector<float>& scales,
                 long long d, long long shard_off, long long shard_len, int rank,
                 const std::vector<uint64_t>& agg, const std::vector<uint64_t>& pads, uint32_t epoch,
                 uint64_t epoch_ptr, uint64_t counter, uint64_t status, const std::vector<uint64_t>& upd_params,
                 const std::vector<uint64_t>& upd_moms, float lr, float mu, float wd, int grid_limit,
                 long long rng_off, long long rng_len, uint32_t seq_mul, uint32_t seq_add, uint64_t agg_mc,
                 uint32_t live_mask, double spin_s, uint64_t trace = 0) {
  std::memset(&a, 0, sizeof(a));
  if (rows.empty() || rows.size() > BZ_MAXN) throw std::invalid_argument("rows");
  if (agg.size() != pads.size() || agg.empty() || agg.size() > BZ_MAXW) throw std::invalid_argument("agg/pads");
  for (size_t i = 0; i < BZ_MAXN; ++i) {
    a.rows.p[i] = i < rows.size() ? as_ptr<const float>(rows[i]) : nullptr;
    a.scales.s[i] = (i < scales.size()) ? scales[i] : 1.0f;
  }
  a.n = (int)rows.size();
  a.d = d;
  a.shard_off = shard_off;
  a.shard_len = shard_len;
  a.rng_off = rng_off;
  a.rng_len = rng_len;
  a.rank = rank;
  a.world = (int)agg.size();
  a.live_mask = live_mask;
  for (size_t p = 0; p < agg.size(); ++p) {
    a.agg[p] = as_ptr<float>(agg[p]);
    a.pad[p] = as_ptr<uint32_t>(pads[p]);
  }
  a.agg_mc = as_ptr<float>(agg_mc);
  a.epoch = epoch;
  a.epoch_ptr = as_ptr<const uint32_t>(epoch_ptr);
  a.seq_mul = seq_mul;
  a.seq_add = seq_add;
  a.spin_ns = (unsigned long long)(spin_s * 1e9);
  a.trace = as_ptr<unsigned long long>(trace);
  a.counter = as_ptr<unsigned int>(counter);
  a.status = as_ptr<int>(status);
  if (upd_params.size() > BZ_MAXR) throw std::invalid_argument("too many replicas");
  a.upd.count = (int)upd_params.size();
  for (size_t r = 0; r < upd_params.size(); ++r) {
    a.upd.param[r] = as_ptr<float>(upd_params[r]);
    a.upd.mom[r] = upd_moms.empty() ? nullptr : as_ptr<float>(upd_moms[r]);
  }
  a.upd.lr = lr;
  a.upd.mu = mu;
  a.upd.wd = wd;
  a.grid_limit = grid_limit;
}

}  // namespace

void bz_bind_runtime(py::module_& m) {
  m.def("device_count", [] {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
  });
  m.def("sm_count", [](int device) {
    int v = 0;
    check(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device), "sm_count");
    return v;
  });
  m.def("raw_alloc", [](size_t bytes) {
    void* p = nullptr;
    check(cudaMalloc(&p, bytes), "cudaMalloc");
    check(cudaMemset(p, 0, bytes), "cudaMemset");
    return (uint64_t)(uintptr_t)p;
  });
  m.def("raw_free", [](uint64_t p) { check(cudaFree(as_ptr<void>(p)), "cudaFree"); });
  m.def("ipc_export", [](uint64_t p) {
    cudaIpcMemHandle_t h;
    check(cudaIpcGetMemHandle(&h, as_ptr<void>(p)), "cudaIpcGetMemHandle");
    return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
  });
  m.def("ipc_open", [](const std::string& handle) {
    if (handle.size() != sizeof(cudaIpcMemHandle_t)) throw std::invalid_argument("bad IPC handle");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle.data(), sizeof(h));
    void* p = nullptr;
    check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return (uint64_t)(uintptr_t)p;
  });
  m.def("ipc_close", [](uint64_t p) { check(cudaIpcCloseMemHandle(as_ptr<void>(p)), "cudaIpcCloseMemHandle"); });
  m.def("can_access_peer", [](int dev, int peer) {
    int ok = 0;
    check(cudaDeviceCanAccessPeer(&ok, dev, peer), "cudaDeviceCanAccessPeer");
    return ok != 0;
  });
  m.def("enable_peer_access", [](int peer) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
      return;
    }
    check(e, "cudaDeviceEnablePeerAccess");
  });
  m.def("memset32_async", [](uint64_t p, uint32_t value, size_t words, uint64_t stream) {
    if (value == 0) {
      check(cudaMemsetAsync(as_ptr<void>(p), 0, words * 4, as_stream(stream)), "cudaMemsetAsync");
    } else {
      std::vector<uint32_t> host(words, value);
      check(cudaMemcpyAsync(as_ptr<void>(p), host.data(), words * 4, cudaMemcpyHostToDevice,
                            as_stream(stream)),
            "cudaMemcpyAsync");
      check(cudaStreamSynchronize(as_stream(stream)), "cudaStreamSynchronize");
    }
  });
  m.def("read_i32", [](uint64_t p) {
    int v = 0;
    check(cudaMemcpy(&v, as_ptr<void>(p), 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
    return v;
  });

  m.def(
      "fused_ps_wsum",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, uint64_t W, long long d,
         long long shard_off, long long shard_len, int rank, const std::vector<uint64_t>& agg,
         const std::vector<uint64_t>& pads, uint64_t epoch_ptr, uint64_t counter, uint64_t status,
         const std::vector<uint64_t>& upd_params, const std::vector<uint64_t>& upd_moms, float lr,
         float mu, float wd, int sm_count, uint64_t stream, int grid_limit, long long rng_off,
         long long rng_len, uint32_t seq_mul, uint32_t seq_add, uint64_t agg_mc, uint32_t live_mask,
         double spin_s) {
        BzFusedPsArgs a;
        fill_common(a, rows, scales, d, shard_off, shard_len, rank, agg, pads, 0u, epoch_ptr, counter, status,
                    upd_params, upd_moms, lr, mu, wd, grid_limit, rng_off, rng_len, seq_mul, seq_add, agg_mc,
                    live_mask, spin_s);
        a.W = as_ptr<const float>(W);
        int e = bz_fused_ps_wsum(&a, sm_count, as_stream(stream));
        if (e != 0)
          throw std::runtime_error(std::string("fused_ps_wsum: CUDA error ") +
                                   cudaGetErrorString((cudaError_t)e));
      },
      py::arg("rows"), py::arg("scales"), py::arg("W"), py::arg("d"), py::arg("shard_off"),
      py::arg("shard_len"), py::arg("rank"), py::arg("agg"), py::arg("pads"), py::arg("epoch_ptr"),
      py::arg("counter"), py::arg("status"), py::arg("upd_params"), py::arg("upd_moms"),
      py::arg("lr"), py::arg("mu"), py::arg("wd"), py::arg("sm_count"), py::arg("stream"),
      py::arg("grid_limit") = 0, py::arg("rng_off") = 0, py::arg("rng_len") = 0, py::arg("seq_mul") = 0,
      py::arg("seq_add") = 0, py::arg("agg_mc") = 0, py::arg("live_mask") = 0, py::arg("spin_s") = 0.0);

  m.def(
      "flag_barrier",
      [](const std::vector<uint64_t>& pads, int rank, int slot, uint64_t epoch_ptr, uint64_t status,
         uint64_t stream, uint32_t seq_mul, uint32_t seq_add, uint32_t live_mask, double spin_s) {
        BzFlagBarrierArgs a;
        std::memset(&a, 0, sizeof(a));
        if (pads.empty() || pads.size() > BZ_MAXW) throw std::invalid_argument("pads");
        a.world = (int)pads.size();
        a.rank = rank;
        a.slot = slot;
        for (size_t p = 0; p < pads.size(); ++p) a.pad[p] = as_ptr<uint32_t>(pads[p]);
        a.epoch_ptr = as_ptr<const uint32_t>(epoch_ptr);
        a.status = as_ptr<int>(status);
        a.seq_mul = seq_mul;
        a.seq_add = seq_add;
        a.live_mask = live_mask;
        a.spin_ns = (unsigned long long)(spin_s * 1e9);
        int e = bz_flag_barrier(&a, as_stream(stream));
        if (e != 0) throw std::runtime_error("flag_barrier failed");
      },
      py::arg("pads"), py::arg("rank"), py::arg("slot"), py::arg("epoch_ptr"), py::arg("status"),
      py::arg("stream"), py::arg("seq_mul") = 0, py::arg("seq_add") = 0, py::arg("live_mask") = 0,
      py::arg("spin_s") = 0.0);

  m.def("gram_exchange", [](uint64_t local, const std::vector<uint64_t>& slots,
                            const std::vector<uint64_t>& pads, int rank, int n, uint64_t epoch_ptr,
                            uint64_t status, uint64_t out64, uint64_t out32, uint64_t stream,
                            uint32_t live_mask, double spin_s, uint64_t slots_mc) {
    BzGramExchangeArgs a;
    std::memset(&a, 0, sizeof(a));
    if (slots.size() != pads.size() || slots.empty() || slots.size() > BZ_MAXW)
      throw std::invalid_argument("slots/pads");
    a.local = as_ptr<const double>(local);
    a.world = (int)slots.size();
    a.rank = rank;
    a.n = n;
    for (size_t p = 0; p < slots.size(); ++p) {
      a.slots[p] = as_ptr<double>(slots[p]);
      a.pad[p] = as_ptr<uint32_t>(pads[p]);
    }
    a.epoch_ptr = as_ptr<const uint32_t>(epoch_ptr);
    a.status = as_ptr<int>(status);
    a.out64 = as_ptr<double>(out64);
    a.out32 = as_ptr<float>(out32);
    a.live_mask = live_mask;
    a.spin_ns = (unsigned long long)(spin_s * 1e9);
    a.slots_mc = as_ptr<double>(slots_mc);
    int e = bz_gram_exchange(&a, as_stream(stream));
    if (e != 0) throw std::runtime_error("gram_exchange failed");
  }, py::arg("local"), py::arg("slots"), py::arg("pads"), py::arg("rank"), py::arg("n"), py::arg("epoch_ptr"),
     py::arg("status"), py::arg("out64"), py::arg("out32"), py::arg("stream"), py::arg("live_mask") = 0,
     py::arg("spin_s") = 0.0, py::arg("slots_mc") = 0);
  m.attr("PAD_READY") = BZ_PAD_READY;
  m.attr("PAD_DONE") = BZ_PAD_DONE;
  m.attr("PAD_GRAM") = BZ_PAD_GRAM;

  m.def("stamp", [](uint64_t p, uint64_t stream) {
    if (bz_stamp(as_ptr<unsigned long long>(p), as_stream(stream)) != 0) throw std::runtime_error("stamp failed");
  });
  m.def("bump_u32", [](uint64_t p, uint64_t stream) {
    int e = bz_bump_u32(as_ptr<uint32_t>(p), as_stream(stream));
    if (e != 0) throw std::runtime_error("bump_u32 failed");
  });
  m.attr("PAD_WORDS") = BZ_PAD_WORDS;
  m.attr("MAXW") = BZ_MAXW;

  m.def(
      "fused_ps_cw",
      [](const std::vector<uint64_t>& rows, const std::vector<float>& scales, int mode, int f,
         int n_virtual, int n_honest, float va, float vb, long long d, long long shard_off,
         long long shard_len, int rank, const std::vector<uint64_t>& agg,
         const std::vector<uint64_t>& pads, uint32_t epoch, uint64_t epoch_ptr, uint64_t counter, uint64_t status,
         const std::vector<uint64_t>& upd_params, const std::vector<uint64_t>& upd_moms, float lr,
         float mu, float wd, int sm_count, uint64_t stream, int grid_limit, long long rng_off,
         long long rng_len, uint32_t seq_mul, uint32_t seq_add, uint64_t agg_mc, uint32_t live_mask,
         double spin_s, uint64_t trace) {
        BzFusedPsArgs a;
        fill_common(a, rows, scales, d, shard_off, shard_len, rank, agg, pads, epoch, epoch_ptr, counter, status,
                    upd_params, upd_moms, lr, mu, wd, grid_limit, rng_off, rng_len, seq_mul, seq_add, agg_mc,
                    live_mask, spin_s, trace);
        a.virt.count = n_virtual;
        a.virt.n_honest = n_honest;
        a.virt.a = va;
        a.virt.b = vb;
        a.f = f;
        a.mode = mode;
        int e = bz_fused_ps_cw(&a, sm_count, as_stream(stream));
        if (e != 0)
          throw std::runtime_error(std::string("fused_ps_cw: CUDA error ") +
                                   cudaGetErrorString((cudaError_t)e));
      },
      py::arg("rows"), py::arg("scales"), py::arg("mode"), py::arg("f"), py::arg("n_virtual"),
      py::arg("n_honest"), py::arg("va"), py::arg("vb"), py::arg("d"), py::arg("shard_off"),
      py::arg("shard_len"), py::arg("rank"), py::arg("agg"), py::arg("pads"), py::arg("epoch"), py::arg("epoch_ptr"),
      py::arg("counter"), py::arg("status"), py::arg("upd_params"), py::arg("upd_moms"),
      py::arg("lr"), py::arg("mu"), py::arg("wd"), py::arg("sm_count"), py::arg("stream"),
      py::arg("grid_limit") = 0, py::arg("rng_off") = 0, py::arg("rng_len") = 0, py::arg("seq_mul") = 0,
      py::arg("seq_add") = 0, py::arg("agg_mc") = 0, py::arg("live_mask") = 0, py::arg("spin_s") = 0.0,
      py::arg("trace") = 0);
}
