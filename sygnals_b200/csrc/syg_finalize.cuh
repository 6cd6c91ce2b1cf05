// sygnals_b200/csrc/syg_finalize.cuh -- per-unit epilogue (included by exactly one translation unit)
#pragma once

#include "syg_device.cuh"
#include "syg_params.h"
#include "syg_finalize_dev.cuh"

namespace sygdev {

// --------------------------------------------------------------------------------------------------------
// finalize: grid = (ceil(T/TT), n_units), kThreads threads.
//   MFCC: S_db = 10 log10(max(amin, mel)) - 10 log10(max(amin, max_unit mel)); S_db = max(S_db, max(S_db) - top_db)
//         (librosa.power_to_db(ref=np.max), manager.py:223) then DCT rows (cepstral.py:106-115).
//   contrast: power_to_db(peak) - power_to_db(valley), each clamped to its own unit-wide max - top_db
//         (librosa.feature.spectral_contrast, frequency_domain.py:200-207).
// --------------------------------------------------------------------------------------------------------

// Persistent: the chunk's frames are one flat sequence cut into tiles of kFinTT consecutive frames (a tile may straddle units:
// every slot carries its own unit, reference level and clamps), CTAs stride over the tiles.  Short units (T = 101 in the
// speech-commands shape) no longer leave partial tiles or one tiny CTA per 32 frames.
template <int kFinTT>
__global__ void __launch_bounds__(kThreads) finalize_kernel_t(const syg::FinalizeArgs a) {
    SYG_DYN_SMEM(smem_raw);
    __shared__ long long s_out[kFinTT];         // offset of (unit, row 0, frame t) in `out`
    __shared__ float s_ref[kFinTT], s_floor[kFinTT], s_pfloor[kFinTT], s_vfloor[kFinTT];
    const int tid = threadIdx.x;
    const long long total = a.n_units * (long long)a.T;
    const long long n_tiles = (total + kFinTT - 1) / kFinTT;
    const int N = a.n_mels;
    const int H = a.dct_fold ? (N + 1) / 2 : N;
    const int H4 = (H + 3) / 4 * 4;                                     // DMMA k-steps of 4 columns
    const int P = fin_pitch(N);
    float* const xs = reinterpret_cast<float*>(smem_raw);              // [kFinTT][P] S_db (FP32, as the reference's float32->float64 values)
    const int warp = tid >> 5, lane = tid & 31;

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
            // S_db tile: the chunk's raw mel energies of these frames are ONE contiguous block of nt * N floats -> coalesced 16-byte
            // loads, four in flight per thread before the first use (the kernel is bound by the latency of these loads), dB
            // conversion with the frame's reference / clamp, FP32 tile in shared memory.
            {
                const float* const blk = a.melws + gf0 * N;             // 128 N bytes per full tile: 16-byte aligned
                const int n_el = nt * N, n_ch = (n_el + 3) >> 2;
                for (int c0 = tid; c0 < n_ch; c0 += 4 * kThreads) {
                    float4 v[4];
                    SYG_UNROLL
                    for (int r = 0; r < 4; ++r) {
                        const int c = c0 + r * kThreads, e = 4 * c;
                        v[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (e + 3 < n_el) v[r] = __ldg(reinterpret_cast<const float4*>(blk) + c);
                        else if (e < n_el) {
                            v[r].x = __ldg(blk + e);
                            if (e + 1 < n_el) v[r].y = __ldg(blk + e + 1);
                            if (e + 2 < n_el) v[r].z = __ldg(blk + e + 2);
                        }
                    }
                    SYG_UNROLL
                    for (int r = 0; r < 4; ++r) {
                        const int e = 4 * (c0 + r * kThreads);
                        if (e >= n_el) break;
                        int tt = e / N, n = e - tt * N;
                        const float in[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
                        if (n + 3 < N && e + 3 < n_el) {                // chunk inside one row (always, when N is a multiple of 4)
                            const float ref_db = s_ref[tt], floor_db = s_floor[tt];
                            float4 o;
                            o.x = fmaxf(db10(fmaxf(a.amin, in[0])) - ref_db, floor_db);
                            o.y = fmaxf(db10(fmaxf(a.amin, in[1])) - ref_db, floor_db);
                            o.z = fmaxf(db10(fmaxf(a.amin, in[2])) - ref_db, floor_db);
                            o.w = fmaxf(db10(fmaxf(a.amin, in[3])) - ref_db, floor_db);
                            if ((n & 3) == 0) *reinterpret_cast<float4*>(xs + tt * P + n) = o;
                            else { xs[tt * P + n] = o.x; xs[tt * P + n + 1] = o.y; xs[tt * P + n + 2] = o.z; xs[tt * P + n + 3] = o.w; }
                        } else {
                            for (int q = 0; q < 4 && e + q < n_el; ++q) {
                                xs[tt * P + n] = fmaxf(db10(fmaxf(a.amin, in[q])) - s_ref[tt], s_floor[tt]);
                                if (++n == N) { n = 0; ++tt; }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // DCT as FP64 tensor-core products (mma.sync m8n8k4, DMMA): D[coefficient][frame] += dct[coefficient][n] * S[frame][n].
            // Warp w owns one parity (even coefficients see the folded sums, odd ones the differences: see the B fragment below) and one
            // n-tile of 8 frames; 8 coefficients of that parity per m-tile.  Per k-step of 4 mel columns a lane loads ONE table
            // entry and ONE tile entry for 256 multiply-adds of the warp (the scalar version needed 1.5 loads per FMA and was
            // bound by the load/store unit: 3.5 ms per 10 h of audio).
            {
                const int par = warp & 1;                               // kThreads / 32 = 8 warps: 2 parities x 4 n-tiles of 8 frames (per 32 frames)
                const int g = lane >> 2, q = lane & 3;
                const int n_par = par ? a.n_mfcc / 2 : (a.n_mfcc + 1) / 2;
                const double sgn = par ? -1.0 : 1.0;
                for (int nt8 = (warp >> 1) * 8; nt8 < kFinTT; nt8 += 32)
                for (int m0 = 0; m0 < n_par; m0 += 8) {
                    const float* const xrow = xs + (nt8 + g) * P;        // this lane's frame (B column)
                    const int c_a = par + 2 * (m0 + g);                 // coefficient of this lane's A row
                    const bool a_ok = c_a < a.n_mfcc;
                    const double* const drow = a.dct + (long long)(a_ok ? c_a : 0) * N + q;
                    double d0 = 0.0, d1 = 0.0;
                    for (int k0 = 0; k0 < H4; k0 += 4) {
                        const int k = k0 + q;
                        const double av = (a_ok && k < H) ? __ldg(drow + k0) : 0.0;
                        // DCT-II rows are (-1)^c symmetric about the centre (cos(pi c (2(N-1-n)+1) / 2N) = (-1)^c cos(pi c (2n+1) / 2N)):
                        // even coefficients see s[n] + s[N-1-n], odd ones s[n] - s[N-1-n], n < ceil(N/2) (centre of an odd N once);
                        // folded here, in FP64, on the way into the fragment.  Other DCT types read the unfolded row.
                        double bv = 0.0;
                        if (k < H) {
                            bv = (double)xrow[k];
                            if (a.dct_fold) {
                                const int k2 = N - 1 - k;
                                bv = (k2 != k) ? bv + sgn * (double)xrow[k2] : (par ? 0.0 : bv);
                            }
                        }
                        mma_m8n8k4_f64(d0, d1, av, bv);
                    }
                    // lane holds D[row g][cols 2q, 2q+1] = coefficient par + 2 (m0 + g) of frames nt8 + 2q, nt8 + 2q + 1
                    if (a_ok) {
                        const int t0 = nt8 + 2 * q;
                        if (t0 < nt) a.out[s_out[t0] + (long long)(a.row_mfcc + c_a) * a.T] = (float)d0;
                        if (t0 + 1 < nt) a.out[s_out[t0 + 1] + (long long)(a.row_mfcc + c_a) * a.T] = (float)d1;
                    }
                }
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
