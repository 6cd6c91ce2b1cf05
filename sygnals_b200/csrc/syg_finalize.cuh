// sygnals_b200/csrc/syg_finalize.cuh -- per-unit epilogue (included by exactly one translation unit)
#pragma once

#include "syg_device.cuh"
#include "syg_params.h"

namespace sygdev {

// --------------------------------------------------------------------------------------------------------
// finalize: grid = (ceil(T/TT), n_units), kThreads threads.
//   MFCC: S_db = 10 log10(max(amin, mel)) - 10 log10(max(amin, max_unit mel)); S_db = max(S_db, max(S_db) - top_db)
//         (librosa.power_to_db(ref=np.max), manager.py:223) then DCT rows (cepstral.py:106-115).
//   contrast: power_to_db(peak) - power_to_db(valley), each clamped to its own unit-wide max - top_db
//         (librosa.feature.spectral_contrast, frequency_domain.py:200-207).
// --------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) finalize_kernel(const syg::FinalizeArgs a) {
    SYG_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x;
    const long long u = blockIdx.y;
    const int t0 = blockIdx.x * kFinTT;
    const int nt = min(kFinTT, a.T - t0);
    const unsigned* um = a.unit_max + u * 4;
    float* const obase = a.out + u * (long long)a.n_rows * a.T;

    if (a.row_mfcc >= 0) {
        // S_db tile in float64: the DCT accumulates in FP64 (|sum| reaches 80*sqrt(n_mels) ~ 905 and the parity bar is 1e-3
        // absolute); converting each S_db value once here (not once per DCT row) keeps the conversion pipe out of the way
        double* const sdb = reinterpret_cast<double*>(smem_raw);            // [kFinTT][n_mels + 1]
        const int ld = a.n_mels + 1;
        const float ref = fmaxf(a.amin, __uint_as_float(um[0]));
        const float ref_db = 10.0f * log10f(ref);
        // the maximum of S_db over the unit is attained at the maximum energy
        const float max_db = 10.0f * log10f(fmaxf(a.amin, __uint_as_float(um[0]))) - ref_db;
        const float floor_db = max_db - a.top_db;
        for (int i = tid; i < nt * a.n_mels; i += kThreads) {
            const int tt = i / a.n_mels, m = i - tt * a.n_mels;
            const float e = a.melws[((u * a.T) + t0 + tt) * a.n_mels + m];
            float db = 10.0f * log10f(fmaxf(a.amin, e)) - ref_db;
            sdb[tt * ld + m] = (double)fmaxf(db, floor_db);
        }
        __syncthreads();
        for (int i = tid; i < a.n_mfcc * kFinTT; i += kThreads) {
            const int c = i / kFinTT, tt = i - c * kFinTT;
            if (tt < nt) {
                const double* d = a.dct + c * a.n_mels;
                const double* s = sdb + tt * ld;
                // four independent chains: a single dependent DFMA chain of n_mels links is latency bound
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                int m = 0;
                for (; m + 4 <= a.n_mels; m += 4) {
                    acc0 = fma(__ldg(&d[m]), s[m], acc0);
                    acc1 = fma(__ldg(&d[m + 1]), s[m + 1], acc1);
                    acc2 = fma(__ldg(&d[m + 2]), s[m + 2], acc2);
                    acc3 = fma(__ldg(&d[m + 3]), s[m + 3], acc3);
                }
                for (; m < a.n_mels; ++m) acc0 = fma(__ldg(&d[m]), s[m], acc0);
                obase[(long long)(a.row_mfcc + c) * a.T + t0 + tt] = (float)((acc0 + acc1) + (acc2 + acc3));
            }
        }
    }
    if (a.nb > 0) {
        const float pmax_db = 10.0f * log10f(fmaxf(a.amin, __uint_as_float(um[1])));
        const float vmax_db = 10.0f * log10f(fmaxf(a.amin, __uint_as_float(um[2])));
        for (int i = tid; i < a.nb * kFinTT; i += kThreads) {
            const int bd = i / kFinTT, tt = i - bd * kFinTT;
            if (tt < nt) {
                const float* c = a.cws + ((u * a.T) + t0 + tt) * (2 * a.nb);
                const float pdb = fmaxf(10.0f * log10f(fmaxf(a.amin, c[bd])), pmax_db - a.top_db);
                const float vdb = fmaxf(10.0f * log10f(fmaxf(a.amin, c[a.nb + bd])), vmax_db - a.top_db);
                obase[(long long)(a.row_contrast + bd) * a.T + t0 + tt] = pdb - vdb;
            }
        }
    }
}

}  // namespace sygdev
