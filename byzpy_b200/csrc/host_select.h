// Host (CPU) coordinate-wise selection networks: see host_select.cpp.
#pragma once
#include <cstdint>

enum { BZ_HOST_MEDIAN = 0, BZ_HOST_TRMEAN = 1, BZ_HOST_MEAMED = 2, BZ_HOST_MEAN = 3 };

// rows: n pointers to d contiguous fp32 values each; scales: n per-row multipliers or nullptr.
// Returns 0 on success, 1 bad shape, 2 bad f, 3 bad mode.
int bz_host_cw_select(const float* const* rows, const float* scales, int n, int64_t d, int mode, int f,
                      float* out, int threads);
// number of comparators of the merge-exchange network for n rows (-1 if n is out of range)
int bz_host_network_size(int n);
// out[j] = a * mean_i(x_ij) + b * population_std_i(x_ij); returns 0 on success.
int bz_host_colstat(const float* const* rows, const float* scales, int n, int64_t d, double a, double b,
                    float* out, int threads);
