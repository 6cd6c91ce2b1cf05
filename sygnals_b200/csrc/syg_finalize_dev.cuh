// sygnals_b200/csrc/syg_finalize_dev.cuh -- device helpers shared by finalize_kernel and the unit-resident stage of the warp kernel
#pragma once

#include "syg_device.cuh"

namespace sygdev {

// 10 log10(x) for x >= amin > 0 through the hardware log2 (absolute error of lg2.approx ~2^-22 -> ~1e-6 dB)
SYG_DEVICE SYG_INLINE float db10(float x) {
#ifdef SYG_EMU
    return 10.0f * log10f(x);
#else
    return 3.0102999566398120f * __log2f(x);
#endif
}

// D (8x8) += A (8x4, row major) * B (4x8, column major) on the FP64 tensor cores.  Fragments (PTX ISA, mma.m8n8k4 .f64), with
// g = lane / 4 and q = lane % 4: a = A[g][q], b = B[q][g], {d0, d1} = D[g][2q], D[g][2q + 1].
SYG_DEVICE SYG_INLINE void mma_m8n8k4_f64(double& d0, double& d1, double av, double bv) {
#ifdef SYG_EMU
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    for (int k = 0; k < 4; ++k) {
        const double ak = __shfl_sync(kFull, av, g * 4 + k);
        const double b0 = __shfl_sync(kFull, bv, (2 * q) * 4 + k), b1 = __shfl_sync(kFull, bv, (2 * q + 1) * 4 + k);
        d0 = fma(ak, b0, d0);
        d1 = fma(ak, b1, d1);
    }
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(av), "d"(bv));
#endif
}

}  // namespace sygdev
