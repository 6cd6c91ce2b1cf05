// sygnals_b200/csrc/syg_welch_warp.cuh
//
// welch_warp_kernel<TL>: Welch / periodogram PSD for nfft <= 2048, warp-synchronous like the feature kernel.
//
// One WARP owns one unit (e.g. one channel-second): it walks the unit's sub-segments FW at a time (FW = 32 / G frames per
// warp), per sub-segment: detrend('constant') -> window -> packed real FFT in registers (FP32x2 butterflies) -> |X|^2 added to
// the warp's private accumulator in shared memory.  After the last sub-segment the FW accumulators are summed in a fixed
// order, scaled (1 / (fs sum w^2) or 1 / (sum w)^2, one-sided doubling except DC / Nyquist) and stored; the unit's rms / crest /
// peak come from one more streaming pass over its samples.  No CTA barrier, no atomics, deterministic.
//
// Reference semantics: scipy.signal.welch / periodogram behind sygnals/core/dsp.py:495-560, :434-493; crest_factor of
// sygnals/core/features/time_domain.py:149-184 applied to the unit.
#pragma once

#include "syg_frame_warp.cuh"

namespace sygdev {

// TBLW: window, twiddles and split twiddles live in shared memory behind the warps' regions (80 of the 111 loads per lane and
// sub-segment are table reads; as LDS they skip the L1 tag stage).  Used where 16 warps + tables fit one CTA (M = 512: 226 KB, with
// the accumulator pitch left unpadded).
template <class TL, int NT, bool TBLW = false>
struct WelchWarpTile {
    using WT = WarpTile<TL, NT>;
    static constexpr int FW = WT::FW, ZS = WT::ZS;
    // accumulator pitch: the FW lane groups of a warp add to their own accumulators in the same 32-bit shared-memory access, so the
    // groups (G = 32 / FW lanes each) must start G banks apart: pitch = G (mod 32), at least WarpTile::PS
    // (WarpTile::PS carries 48 words of slack for the mel sweep of the feature kernel: not needed here, which is what lets the
    // padded pitch fit the 16-warp CTA of the TBLW instantiation)
    static constexpr int PS0 = TBLW ? ((WT::M + 1) + 4 * ((WT::M + 1) >> 5) + 3) / 4 * 4 : WT::PS;
    static constexpr int PS = (FW == 1) ? PS0 : ((PS0 - WT::G + 31) / 32 * 32 + WT::G);
    static constexpr int ZR = (2 * ZS + 3) / 4 * 4;                  // floats of one Z region
    static constexpr int warp_floats = FW * (ZR + PS);               // Z regions, then the accumulators
    static constexpr size_t table_bytes = TBLW ? (size_t)(2 * WT::M * 4 + WT::M * 8 + (WT::M / 2 + 1) * 8) : 0;
    static constexpr size_t bytes = (size_t)WT::kWarps * warp_floats * sizeof(float) + table_bytes;
};

template <class TL, int NT, int MINB, bool TBLW = false>
__global__ void __launch_bounds__(NT, MINB) welch_warp_kernel(const syg::WelchArgs a) {
    using WT = WarpTile<TL, NT>;
    using WW = WelchWarpTile<TL, NT, TBLW>;
    constexpr int E = WT::E, M = WT::M, G = WT::G, FW = WT::FW, R2 = WT::R2, PS = WW::PS, LE = WT::LOG2E;
    constexpr int Q = E / R2;
    static_assert(R2 == G, "two-pass warp tile");
    constexpr bool kShflSplit = (SYG_SPLIT_SHFL != 0);                  // mirrors of the real split by SHFL (syg_device.cuh: mirror_of)
    constexpr bool kHalfZ = TBLW && (SYG_SPLIT_HALF != 0);              // tables in shared memory: Z/2 from pass 2 + split_power_h
#ifndef SYG_WELCH_REGACC
#define SYG_WELCH_REGACC 0
#endif
    // With the shuffle split a lane produces the SAME bins (k = j + G i and M - k) for every sub-segment, so their running sums
    // could stay in registers (E + 1 floats; same additions in the same order: bit-identical).  Measured on B200 (cfg5, nfft 1024):
    // 2.91 ms against 2.78 ms with the shared-memory accumulator -- at 128 registers the extra 33 live values spill.  Off.
    constexpr bool kRegAcc = kShflSplit && (SYG_WELCH_REGACC != 0);
    constexpr int B = M + 1;
    SYG_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int f = lane / G, j = lane % G;
    float* const wbase = reinterpret_cast<float*>(smem_raw) + warp * WW::warp_floats;
    float2* const zs = reinterpret_cast<float2*>(wbase + f * WW::ZR);
    float* const accw = wbase + FW * WW::ZR;                          // [FW][PS], ppad layout
    float* const acc = accw + f * PS;
    float2* const tbw = reinterpret_cast<float2*>(reinterpret_cast<float*>(smem_raw) + WT::kWarps * WW::warp_floats);   // [M] window, [M] tw, [M/2+1] tws
    const float2* const w2 = TBLW ? tbw : reinterpret_cast<const float2*>(a.window);
    const float2* const t_tw = TBLW ? tbw + M : a.tw;
    const float2* const t_tws = TBLW ? tbw + 2 * M : a.tws;
    if (TBLW) {
        for (int i = tid; i < M; i += NT) {
            tbw[i] = __ldg(reinterpret_cast<const float2*>(a.window) + i);
            const float2 w = __ldg(a.tw + (i / E) * (i % E));            // transposed for pass 2: entry [r][k] = W_M^{r k} (as frame_warp_kernel)
            tbw[M + i] = kHalfZ ? make_float2(0.5f * w.x, 0.5f * w.y) : w;
        }
        for (int i = tid; i <= M / 2; i += NT) {                        // split twiddles: tangent form (split_power_h) or with the halving folded in (split_power)
            const float2 w = __ldg(a.tws + i);
            tbw[2 * M + i] = kHalfZ ? split_twiddle_h(w, 4 * i < M) : make_float2(0.5f * w.x, 0.5f * w.y);
        }
        __syncthreads();
    }
    const bool full = (a.nperseg == 2 * M);
    const float inv_n = 1.0f / (float)a.nperseg;

    const long long warps = (long long)gridDim.x * WT::kWarps;
    for (long long u = (long long)blockIdx.x * WT::kWarps + warp; u < a.g.n_units; u += warps) {
        const UnitRef ur = unit_ref(a.g, u);
        const float* yb = a.y + ur.start;
        float ak[E / 2 + 1], am[E / 2];                                // kRegAcc: sums of |X[j + G i]|^2 and of the mirrors |X[M - j - G i]|^2
        if constexpr (kRegAcc) {
            SYG_UNROLL
            for (int i = 0; i <= E / 2; ++i) ak[i] = 0.0f;
            SYG_UNROLL
            for (int i = 0; i < E / 2; ++i) am[i] = 0.0f;
        } else {
            for (int i = lane; i < FW * PS; i += 32) accw[i] = 0.0f;
            __syncwarp();
        }
        // unit rms / crest / peak ride on the sub-segment loads: sub-segment s accounts for its first `step` samples (the rest
        // belongs to its successors), the last one for all of its nperseg samples; samples behind the last sub-segment are
        // swept afterwards.  Every sample of the unit is counted exactly once and read from memory for the FFT only.
        double sqd = 0.0;
        float pk = 0.0f;
        for (int s0 = 0; s0 < a.nseg; s0 += FW) {
            const int s = s0 + f;
            const bool valid = s < a.nseg;
            const long long p0 = (long long)s * a.step;
            const long long nv = valid ? ur.valid : 0;
            // ---------------- load, detrend, window ----------------
            float2 z[E];
            const float* src = yb + p0;
            const bool interior = full && valid && (p0 + 2 * M <= nv) && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0);
            float2 sm01 = make_float2(0.0f, 0.0f), sm23 = sm01;
            if (__all_sync(kFull, interior)) {
                SYG_UNROLL
                for (int r = 0; r < E; ++r) {
                    z[r] = __ldg(reinterpret_cast<const float2*>(src) + j + r * G);
                    if (r & 1) sm23 = __fadd2_rn(sm23, z[r]); else sm01 = __fadd2_rn(sm01, z[r]);
                }
            } else {
                SYG_UNROLL
                for (int r = 0; r < E; ++r) {
                    const int c = j + r * G;
                    const long long pos = p0 + 2 * c;
                    float2 v = make_float2(0.0f, 0.0f);
                    if (2 * c < a.nperseg && pos < nv) v.x = __ldg(yb + pos);
                    if (2 * c + 1 < a.nperseg && pos + 1 < nv) v.y = __ldg(yb + pos + 1);
                    z[r] = v;
                    if (r & 1) sm23 = __fadd2_rn(sm23, v); else sm01 = __fadd2_rn(sm01, v);
                }
            }
            if (a.stats && valid) {
                const int lim = (s == a.nseg - 1) ? a.nperseg : a.step;
                float2 sq = make_float2(0.0f, 0.0f);
                SYG_UNROLL
                for (int r = 0; r < E; ++r) {
                    const int c2 = 2 * (j + r * G);
                    if (c2 + 1 < lim) {
                        sq = __ffma2_rn(z[r], z[r], sq);
                        pk = fmaxf(fmaxf(pk, fabsf(z[r].x)), fabsf(z[r].y));   // one FMNMX3
                    } else if (c2 < lim) {
                        sq.x = __fmaf_rn(z[r].x, z[r].x, sq.x);
                        pk = fmaxf(pk, fabsf(z[r].x));
                    }
                }
                sqd += (double)(sq.x + sq.y);                          // float32 partial sums of <= 2 E terms, float64 across
            }
            float mean = 0.0f;
            if (a.detrend) {                                          // scipy detrend('constant'): subtract the sub-segment mean
                const double tot = lanes_sum<G>((double)((sm01.x + sm01.y) + (sm23.x + sm23.y)));
                mean = (float)(tot * (double)inv_n);
            }
            const float2 nmean2 = make_float2(-mean, -mean);
            SYG_UNROLL
            for (int r = 0; r < E; ++r) {
                const float2 w = TBLW ? w2[j + r * G] : __ldg(w2 + j + r * G);   // zero beyond nperseg: padding samples vanish
                z[r] = __fmul2_rn(__fadd2_rn(z[r], nmean2), w);         // (z - mean) * w, both halves at once
            }
            // ---------------- FFT (as in frame_warp_kernel) ----------------
            dft_dif_p<E, 1>(z);
            SYG_UNROLL
            for (int kp = 0; kp < E; ++kp) zs[zpad<LE>(j * E + kp)] = z[bitrev(kp, LE)];
            __syncwarp();
            SYG_UNROLL
            for (int q = 0; q < Q; ++q) {
                const int b = j + q * G;
                SYG_UNROLL
                for (int r = 0; r < R2; ++r) z[q * R2 + r] = zs[zpad<LE>(b + r * (M / R2))];
            }
            __syncwarp();
            SYG_UNROLL
            for (int q = 0; q < Q; ++q) {
                const int b = j + q * G;
                const int k = b & (E - 1);
                SYG_UNROLL
                for (int r = 1; r < R2; ++r) {
                    const float2 w = TBLW ? t_tw[r * E + k] : __ldg(&t_tw[r * k]);
                    cmul(z[q * R2 + r].x, z[q * R2 + r].y, w.x, w.y);
                }
                dft_dif_p<R2, 1, kHalfZ>(z + q * R2);
                if constexpr (!kShflSplit) {
                    const int ob = (b - k) * R2 + k;
                    SYG_UNROLL
                    for (int kp = 0; kp < R2; ++kp) zs[zpad<LE>(ob + kp * E)] = z[q * R2 + bitrev(kp, ilog2(R2))];
                }
            }
            if constexpr (!kShflSplit) __syncwarp();
            // ---------------- real split -> accumulate |X[k]|^2 ----------------
            {
                const int jz = (j == 0) ? 1 : 0;
                const float2* const zk0 = zs + j;
                const float2* const zm0 = zs - j;
                const float2* const zm1 = zm0 + jz;
                float* const pk0 = acc + j;
                float* const pm0 = acc - j;
                float* const pm1 = pm0 + 4 * jz;
                SYG_UNROLL
                for (int i = 0; i <= E / 2; ++i) {
                    const int kk = i * G;
                    const int k = j + kk;
                    if (i == E / 2 && j != 0) break;
                    float2 zk, zm;
                    if constexpr (kShflSplit) {                        // mirrors from the partner lane's registers (mirror_of)
                        zk = z[zreg_of<E, G>(i)];
                        zm = zk;                                       // i = E/2 (lane 0): bin M/2 pairs with itself
                        if (i < E / 2) zm = mirror_of<E, G>(z, i, j);
                    } else {
                        zk = zk0[kk + (kk >> LE)];
                        const int c1 = (M - kk) + ((M - kk - 1) >> LE);
                        const bool blk = ((M - kk) & (E - 1)) == 0;
                        zm = blk ? zm1[c1] : zm0[c1];
                        if (i == 0 && j == 0) zm = zk;
                    }
                    float2 w = TBLW ? t_tws[k] : __ldg(&t_tws[k]);
                    if (!TBLW) w = make_float2(0.5f * w.x, 0.5f * w.y);
                    float pwk, pwm;
                    if constexpr (kHalfZ) split_power_h(4 * i < E, zk, zm, w, pwk, pwm);
                    else split_power(zk, zm, w, pwk, pwm);             // |X[k]|^2, |X[M-k]|^2 straight from the packed pair (14 instead of 22 operations)
                    if constexpr (kRegAcc) {
                        ak[i] += valid ? pwk : 0.0f;
                        if (i < E / 2) am[i] += valid ? pwm : 0.0f;
                    } else if (valid) {
                        pk0[kk + ((kk >> 5) << 2)] += pwk;
                        const int q1 = (M - kk) + (((M - kk - 1) >> 5) << 2);
                        const bool blk5 = ((M - kk) & 31) == 0;
                        if (2 * k != M) (blk5 ? pm1 : pm0)[q1] += pwm;
                    }
                }
            }
            __syncwarp();
        }
        if constexpr (kRegAcc) {                                        // every bin of the frame group's accumulator is written exactly once
            float* const pk0 = acc + j;
            float* const pm0 = acc - j;
            float* const pm1 = pm0 + 4 * ((j == 0) ? 1 : 0);
            SYG_UNROLL
            for (int i = 0; i <= E / 2; ++i) {
                const int kk = i * G;
                if (i == E / 2 && j != 0) break;
                pk0[kk + ((kk >> 5) << 2)] = ak[i];
                if (i < E / 2) {
                    const int q1 = (M - kk) + (((M - kk - 1) >> 5) << 2);
                    const bool blk5 = ((M - kk) & 31) == 0;
                    (blk5 ? pm1 : pm0)[q1] = am[i];
                }
            }
            __syncwarp();
        }
        // ---------------- mean over sub-segments (fixed order), scaling, store ----------------
        const float inv = a.scale / (float)a.nseg;
        for (int k = lane; k < B; k += 32) {
            float v = 0.0f;
            SYG_UNROLL
            for (int ff = 0; ff < FW; ++ff) v += accw[ff * PS + ppad(k)];
            v *= inv;
            if (a.onesided_double && k != 0 && k != M) v *= 2.0f;
            a.psd[u * B + k] = v;
        }
        if (a.stats) {                                                // rms / crest / peak of the whole unit
            float2 sq = make_float2(0.0f, 0.0f);
            int run = 0;
            const long long covered = a.nseg > 0 ? (long long)(a.nseg - 1) * a.step + a.nperseg : 0;
            for (long long i = covered + lane; i < a.g.unit_len; i += 32) {
                const float v = (i < ur.valid) ? __ldg(yb + i) : 0.0f;
                sq.x = __fmaf_rn(v, v, sq.x);
                pk = fmaxf(pk, fabsf(v));
                if (++run == 64) { sqd += (double)sq.x; sq.x = 0.0f; run = 0; }
            }
            sqd += (double)sq.x;
            SYG_UNROLL
            for (int o = 16; o >= 1; o >>= 1) {
                sqd += __shfl_xor_sync(kFull, sqd, o);
                pk = fmaxf(pk, __shfl_xor_sync(kFull, pk, o));
            }
            if (lane == 0) {
                const double rms = a.g.unit_len > 0 ? sqrt(sqd / (double)a.g.unit_len) : 0.0;
                a.stats[u * 3 + 0] = (float)rms;
                a.stats[u * 3 + 1] = (rms < kEps64) ? 0.0f : (float)((double)pk / rms);
                a.stats[u * 3 + 2] = pk;
            }
        }
        __syncwarp();
    }
}

}  // namespace sygdev
