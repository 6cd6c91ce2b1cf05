// helpers shared by the kernel translation units
#pragma once

#include <algorithm>
#include <string>

#include "syg_launch.h"

#ifndef SYG_EMU
#define SYG_OCCUPANCY(nb, kernel, threads, smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&(nb), kernel, threads, smem)
#else
#define SYG_OCCUPANCY(nb, kernel, threads, smem) ((nb) = 2, cudaSuccess)
#endif

#define LCK(expr)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e_); return -3; } \
    } while (0)

namespace syglaunch {
inline int ilog2i(long long v) { int l = 0; while ((1LL << l) < v) ++l; return l; }
}
