// Native runtime pieces bound into byzpy_b200._C: CUDA-IPC symmetric memory,
// peer access, raw-memory helpers and the fused parameter-server launchers.
#pragma once
#include <pybind11/pybind11.h>
void bz_bind_runtime(pybind11::module_& m);
// symmetric heap on the VMM API + NVLS multicast (vmm.cpp)
void bz_bind_vmm(pybind11::module_& m);
