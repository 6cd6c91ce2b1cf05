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

// 10 log10(x) for x >= amin > 0 through the hardware log2 (absolute error of lg2.approx ~2^-22 -> ~1e-6 dB)
SYG_DEVICE SYG_INLINE float db10(float x) {
#ifdef SYG_EMU
    return 10.0f * log10f(x);
#else
    return 3.0102999566398120f * __log2f(x);
#endif
}

// Persistent: the chunk's frames are one flat sequence cut into tiles of kFinTT consecutive frames (a tile may straddle units:
// every slot carries its own unit, reference level and clamps), CTAs stride over the tiles.  Short units (T = 101 in the
// speech-commands shape) no longer leave partial tiles or one tiny CTA per 32 frames.
__global__ void __launch_bounds__(kThreads) finalize_kernel(const syg::FinalizeArgs a) {
    SYG_DYN_SMEM(smem_raw);
    __shared__ long long s_out[kFinTT];         // offset of (unit, row 0, frame t) in `out`
    __shared__ float s_ref[kFinTT], s_floor[kFinTT], s_pfloor[kFinTT], s_vfloor[kFinTT];
    const int tid = threadIdx.x;
    const long long total = a.n_units * (long long)a.T;
    const long long n_tiles = (total + kFinTT - 1) / kFinTT;
    const int N = a.n_mels;
    const int H = a.dct_fold ? (N + 1) / 2 : N;
    const int ld = H + 1;
    double* const se = reinterpret_cast<double*>(smem_raw);            // [kFinTT][H + 1]
    double* const so = a.dct_fold ? se + kFinTT * ld : se;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long gf0 = tile * kFinTT;
        const int nt = (int)min((long long)kFinTT, total - gf0);
        if (tid < nt) {
            const long long gf = gf0 + tid;
            const long long u = gf / a.T;
            const int t = (int)(gf - u * a.T);
            const unsigned* um = a.unit_max + u * 4;
            s_out[tid] = u * (long long)a.n_rows * a.T + t;
            // power_to_db(ref=np.max): the maximum of S_db over the unit is attained at the maximum energy -> 0 dB
            const float ref_db = db10(fmaxf(a.amin, __uint_as_float(um[0])));
            s_ref[tid] = ref_db;
            s_floor[tid] = -a.top_db;
            s_pfloor[tid] = db10(fmaxf(a.amin, __uint_as_float(um[1]))) - a.top_db;
            s_vfloor[tid] = db10(fmaxf(a.amin, __uint_as_float(um[2]))) - a.top_db;
        }
        __syncthreads();
        if (a.row_mfcc >= 0) {
            // S_db tile in float64: the DCT accumulates in FP64 (|sum| reaches 80*sqrt(n_mels) ~ 905 and the parity bar is 1e-3
            // absolute).  DCT-II rows are (-1)^k symmetric about the centre (cos(pi k (2(N-1-n)+1) / 2N) = (-1)^k cos(pi k (2n+1) / 2N)),
            // so the tile is stored folded: se[n] = s[n] + s[N-1-n], so[n] = s[n] - s[N-1-n] (n < N/2; centre term of an odd N
            // separately) and every coefficient needs H = ceil(N/2) products.  Other DCT types use the unfolded tile (se = so = s).
            for (int i = tid; i < nt * H; i += kThreads) {
                const int tt = i / H, n = i - tt * H;
                const float* row = a.melws + (gf0 + tt) * N;
                const float ref_db = s_ref[tt], floor_db = s_floor[tt];
                const float x = fmaxf(db10(fmaxf(a.amin, row[n])) - ref_db, floor_db);
                if (a.dct_fold) {
                    const int n2 = N - 1 - n;
                    if (n2 != n) {
                        const float y = fmaxf(db10(fmaxf(a.amin, row[n2])) - ref_db, floor_db);
                        se[tt * ld + n] = (double)x + (double)y;
                        so[tt * ld + n] = (double)x - (double)y;
                    } else {
                        se[tt * ld + n] = (double)x;
                        so[tt * ld + n] = 0.0;
                    }
                } else {
                    se[tt * ld + n] = (double)x;
                }
            }
            __syncthreads();
            // work item = (frame, pair of coefficients of equal parity): the pair shares every load of the folded tile
            const int n_ev = (a.n_mfcc + 1) / 2, n_od = a.n_mfcc / 2;
            const int it_ev = (n_ev + 1) / 2, it_od = (n_od + 1) / 2;
            for (int i = tid; i < (it_ev + it_od) * kFinTT; i += kThreads) {
                const int item = i / kFinTT, tt = i - item * kFinTT;
                if (tt >= nt) continue;
                const bool odd = item >= it_ev;
                const int c0 = odd ? 1 + 4 * (item - it_ev) : 4 * item;
                const int c1 = c0 + 2;
                const bool two = c1 < a.n_mfcc;
                const double* d0 = a.dct + (long long)c0 * N;
                const double* d1 = a.dct + (long long)(two ? c1 : c0) * N;
                const double* s = (odd ? so : se) + tt * ld;
                double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;              // two chains per coefficient (DFMA latency)
                int n = 0;
                for (; n + 2 <= H; n += 2) {
                    const double s0 = s[n], s1 = s[n + 1];
                    p0 = fma(__ldg(&d0[n]), s0, p0);
                    q0 = fma(__ldg(&d1[n]), s0, q0);
                    p1 = fma(__ldg(&d0[n + 1]), s1, p1);
                    q1 = fma(__ldg(&d1[n + 1]), s1, q1);
                }
                if (n < H) {
                    p0 = fma(__ldg(&d0[n]), s[n], p0);
                    q0 = fma(__ldg(&d1[n]), s[n], q0);
                }
                float* const o = a.out + s_out[tt];
                o[(long long)(a.row_mfcc + c0) * a.T] = (float)(p0 + p1);
                if (two) o[(long long)(a.row_mfcc + c1) * a.T] = (float)(q0 + q1);
            }
        }
        if (a.nb > 0) {
            for (int i = tid; i < a.nb * kFinTT; i += kThreads) {
                const int bd = i / kFinTT, tt = i - bd * kFinTT;
                if (tt < nt) {
                    const float* c = a.cws + (gf0 + tt) * (2 * a.nb);
                    const float pdb = fmaxf(db10(fmaxf(a.amin, c[bd])), s_pfloor[tt]);
                    const float vdb = fmaxf(db10(fmaxf(a.amin, c[a.nb + bd])), s_vfloor[tt]);
                    a.out[s_out[tt] + (long long)(a.row_contrast + bd) * a.T] = pdb - vdb;
                }
            }
        }
        __syncthreads();                                                // the tile buffers are reused
    }
}

}  // namespace sygdev
