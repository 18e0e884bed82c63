// Symmetric heap on the CUDA virtual-memory-management API with an NVLS multicast alias.
//
// Every rank cuMemCreate's one physical allocation on its GPU (exportable as a POSIX file
// descriptor), maps every peer's allocation into its own address space (unicast P2P views: the
// fused kernels load peer gradient rows through them over NVLink 5 / NVSwitch) and binds its
// allocation to one multicast object shared by the whole team.  The multicast object is mapped a
// second time: a `multimem.st` to that alias is replicated by the switch into every rank's HBM
// (the broadcast half of the fused round is ONE store per tile instead of `world` peer stores), a
// `multimem.ld_reduce` sums the team's copies inside the switch (the (n, n) partial-Gram
// all-reduce).  File descriptors travel between the ranks over Unix sockets
// (parallel/symmetric.py); torch.distributed only carries the socket paths.
//
// The driver library is reached through dlopen so the extension still imports (and every CPU
// test runs) on a box without a GPU driver.  The CUDA-IPC heap in runtime.cpp remains the
// fallback when VMM / POSIX-fd export / multicast are not supported.
//
// Replaces the reference's host shared-memory store + pickled tensor transport (reference
// engine/storage/shared_store.py:21-54, engine/actor/transports/ucx.py:225-270) and the
// fan-out loop of reference engine/parameter_server/ps.py:140-143.
#include <cuda.h>
#include <cuda_runtime.h>

#include "gram_umma.h"
#include <dlfcn.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <unistd.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace py = pybind11;

namespace {

struct Driver {
  void* lib = nullptr;
  bool ok = false;
  std::string err;
#define BZ_DRV(name) decltype(&::name) name = nullptr;
  BZ_DRV(cuInit)
  BZ_DRV(cuDeviceGet)
  BZ_DRV(cuDeviceGetAttribute)
  BZ_DRV(cuGetErrorString)
  BZ_DRV(cuMemCreate)
  BZ_DRV(cuMemRelease)
  BZ_DRV(cuMemAddressReserve)
  BZ_DRV(cuMemAddressFree)
  BZ_DRV(cuMemMap)
  BZ_DRV(cuMemUnmap)
  BZ_DRV(cuMemSetAccess)
  BZ_DRV(cuMemExportToShareableHandle)
  BZ_DRV(cuMemImportFromShareableHandle)
  BZ_DRV(cuMemGetAllocationGranularity)
  BZ_DRV(cuMulticastCreate)
  BZ_DRV(cuMulticastAddDevice)
  BZ_DRV(cuMulticastBindMem)
  BZ_DRV(cuMulticastUnbind)
  BZ_DRV(cuMulticastGetGranularity)
  BZ_DRV(cuTensorMapEncodeTiled)
#undef BZ_DRV
};

Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!d.lib) {
      d.err = "libcuda.so.1 not found";
      return;
    }
    bool all = true;
#define BZ_LOAD(name, sym)                                          \
  d.name = reinterpret_cast<decltype(d.name)>(dlsym(d.lib, sym));   \
  if (!d.name) {                                                    \
    all = false;                                                    \
    d.err += std::string(sym) + " ";                                \
  }
    BZ_LOAD(cuInit, "cuInit")
    BZ_LOAD(cuDeviceGet, "cuDeviceGet")
    BZ_LOAD(cuDeviceGetAttribute, "cuDeviceGetAttribute")
    BZ_LOAD(cuGetErrorString, "cuGetErrorString")
    BZ_LOAD(cuMemCreate, "cuMemCreate")
    BZ_LOAD(cuMemRelease, "cuMemRelease")
    BZ_LOAD(cuMemAddressReserve, "cuMemAddressReserve")
    BZ_LOAD(cuMemAddressFree, "cuMemAddressFree")
    BZ_LOAD(cuMemMap, "cuMemMap")
    BZ_LOAD(cuMemUnmap, "cuMemUnmap")
    BZ_LOAD(cuMemSetAccess, "cuMemSetAccess")
    BZ_LOAD(cuMemExportToShareableHandle, "cuMemExportToShareableHandle")
    BZ_LOAD(cuMemImportFromShareableHandle, "cuMemImportFromShareableHandle")
    BZ_LOAD(cuMemGetAllocationGranularity, "cuMemGetAllocationGranularity")
    BZ_LOAD(cuMulticastCreate, "cuMulticastCreate")
    BZ_LOAD(cuMulticastAddDevice, "cuMulticastAddDevice")
    BZ_LOAD(cuMulticastBindMem, "cuMulticastBindMem")
    BZ_LOAD(cuMulticastUnbind, "cuMulticastUnbind")
    BZ_LOAD(cuMulticastGetGranularity, "cuMulticastGetGranularity")
    BZ_LOAD(cuTensorMapEncodeTiled, "cuTensorMapEncodeTiled")
#undef BZ_LOAD
    if (!all) {
      d.err = "missing driver symbols: " + d.err;
      return;
    }
    if (d.cuInit(0) != CUDA_SUCCESS) {
      d.err = "cuInit failed";
      return;
    }
    d.ok = true;
  });
  return d;
}

Driver& need_driver() {
  Driver& d = driver();
  if (!d.ok) throw std::runtime_error("CUDA driver API unavailable: " + d.err);
  return d;
}

void check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = nullptr;
  Driver& d = driver();
  if (d.cuGetErrorString) d.cuGetErrorString(r, &msg);
  throw std::runtime_error(std::string(what) + ": " + (msg ? msg : "unknown driver error") + " (" +
                           std::to_string((int)r) + ")");
}

// make the primary context of `device` current on this thread (the runtime API owns it)
void bind_device(int device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)
    throw std::runtime_error("cudaSetDevice failed");
}

int attr(int device, CUdevice_attribute a) {
  Driver& d = need_driver();
  CUdevice dev;
  check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
  int v = 0;
  if (d.cuDeviceGetAttribute(&v, a, dev) != CUDA_SUCCESS) return 0;
  return v;
}

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

uint64_t map_handle(CUmemGenericAllocationHandle h, size_t size, size_t align, int device) {
  Driver& d = need_driver();
  CUdeviceptr ptr = 0;
  check(d.cuMemAddressReserve(&ptr, size, align, 0, 0), "cuMemAddressReserve");
  CUresult r = d.cuMemMap(ptr, size, 0, h, 0);
  if (r != CUDA_SUCCESS) {
    d.cuMemAddressFree(ptr, size);
    check(r, "cuMemMap");
  }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.cuMemSetAccess(ptr, size, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.cuMemUnmap(ptr, size);
    d.cuMemAddressFree(ptr, size);
    check(r, "cuMemSetAccess");
  }
  return (uint64_t)ptr;
}

}  // namespace

// 2-D tiled fp32 tensor map for the TMA-fed Gram kernel (gram_umma.cu); see gram_umma.h.
int bz_encode_map_2d(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long cols,
                     unsigned long long row_stride_bytes, unsigned box_rows, unsigned box_cols) {
  Driver& d = driver();
  if (!d.ok) return -1;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)row_stride_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  return (int)d.cuTensorMapEncodeTiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride,
                                       box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

void bz_bind_vmm(py::module_& m) {
  // {"vmm": bool, "posix_fd": bool, "multicast": bool, "reason": str}
  m.def("vmm_support", [](int device) {
    py::dict out;
    Driver& d = driver();
    out["vmm"] = false;
    out["posix_fd"] = false;
    out["multicast"] = false;
    out["reason"] = d.ok ? "" : d.err;
    if (!d.ok) return out;
    try {
      out["vmm"] = attr(device, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED) != 0;
      out["posix_fd"] = attr(device, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED) != 0;
      out["multicast"] = attr(device, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED) != 0;
    } catch (const std::exception& e) {
      out["reason"] = e.what();
    }
    return out;
  });

  // size every rank must round its request up to (allocation and multicast granularities)
  m.def("vmm_granularity", [](int device, int world, bool multicast) {
    Driver& d = need_driver();
    CUmemAllocationProp p = alloc_prop(device);
    size_t g = 0;
    check(d.cuMemGetAllocationGranularity(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    if (multicast) {
      CUmulticastObjectProp mp = {};
      mp.numDevices = (unsigned)world;
      mp.size = g;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t mg = 0;
      check(d.cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM), "cuMulticastGetGranularity");
      if (mg > g) g = mg;
    }
    return g;
  });

  // -> (handle, ptr): zero-filled physical allocation of `size` bytes mapped read-write on `device`
  m.def("vmm_alloc", [](size_t size, size_t align, int device) {
    Driver& d = need_driver();
    bind_device(device);
    CUmemAllocationProp p = alloc_prop(device);
    CUmemGenericAllocationHandle h = 0;
    check(d.cuMemCreate(&h, size, &p, 0), "cuMemCreate");
    uint64_t ptr;
    try {
      ptr = map_handle(h, size, align, device);
    } catch (...) {
      d.cuMemRelease(h);
      throw;
    }
    if (cudaMemset(reinterpret_cast<void*>(ptr), 0, size) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
      throw std::runtime_error("cudaMemset on the VMM allocation failed");
    return py::make_tuple((uint64_t)h, ptr);
  });
  m.def("vmm_export_fd", [](uint64_t handle) {
    Driver& d = need_driver();
    int fd = -1;
    check(d.cuMemExportToShareableHandle(&fd, (CUmemGenericAllocationHandle)handle,
                                         CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
          "cuMemExportToShareableHandle");
    return fd;
  });
  // import a peer's allocation (or the multicast object) from a received fd -> handle
  m.def("vmm_import_fd", [](int fd, int device) {
    Driver& d = need_driver();
    bind_device(device);
    CUmemGenericAllocationHandle h = 0;
    check(d.cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
          "cuMemImportFromShareableHandle");
    return (uint64_t)h;
  });
  m.def("vmm_map", [](uint64_t handle, size_t size, size_t align, int device) {
    bind_device(device);
    return map_handle((CUmemGenericAllocationHandle)handle, size, align, device);
  });
  m.def("vmm_unmap", [](uint64_t ptr, size_t size) {
    Driver& d = need_driver();
    d.cuMemUnmap((CUdeviceptr)ptr, size);
    d.cuMemAddressFree((CUdeviceptr)ptr, size);
  });
  m.def("vmm_release", [](uint64_t handle) { need_driver().cuMemRelease((CUmemGenericAllocationHandle)handle); });
  m.def("close_fd", [](int fd) { ::close(fd); });

  // multicast object: created by one rank, imported (vmm_import_fd) by the others
  m.def("mc_create", [](int world, size_t size) {
    Driver& d = need_driver();
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)world;
    mp.size = size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle h = 0;
    check(d.cuMulticastCreate(&h, &mp), "cuMulticastCreate");
    return (uint64_t)h;
  });
  m.def("mc_add_device", [](uint64_t mc, int device) {
    Driver& d = need_driver();
    CUdevice dev;
    check(d.cuDeviceGet(&dev, device), "cuDeviceGet");
    check(d.cuMulticastAddDevice((CUmemGenericAllocationHandle)mc, dev), "cuMulticastAddDevice");
  });
  // bind this rank's whole allocation at offset 0 of the multicast object (after EVERY rank added its device)
  m.def("mc_bind", [](uint64_t mc, uint64_t mem, size_t size) {
    Driver& d = need_driver();
    check(d.cuMulticastBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem, 0, size, 0),
          "cuMulticastBindMem");
  });
  m.def("mc_unbind", [](uint64_t mc, int device, size_t size) {
    Driver& d = need_driver();
    CUdevice dev;
    if (d.cuDeviceGet(&dev, device) == CUDA_SUCCESS) d.cuMulticastUnbind((CUmemGenericAllocationHandle)mc, dev, 0, size);
  });
}
