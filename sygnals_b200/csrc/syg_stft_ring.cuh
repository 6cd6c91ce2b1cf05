// sygnals_b200/csrc/syg_stft_ring.cuh
//
// stft_ring_kernel<TL, NW, DB, S>: STFT magnitude / power output (compute_stft, sygnals/core/dsp.py:167-229) for n_fft <= 2048
// with the samples staged by the TMA engine.
//
// A CTA of NW warps works in ROUNDS of TT = NW * FW consecutive frames of one unit (FW = frames a warp transforms at once).
//   * input : the samples of a round ((TT-1) hop + n_fft floats, every sample is shared by n_fft/hop frames) arrive in a ring of S
//             shared-memory stages through cp.async.bulk (SASS UBLKCP), one elected thread arming an mbarrier per stage with the
//             byte count; the copy for round i+S is issued as soon as round i's samples are in registers, so the global-load
//             latency of a round is hidden behind the S-1 rounds in front of it and costs no registers (the LDG path of the
//             warp kernel stalled 30 % of its issue slots on the long scoreboard with 16 warps of 128 registers).
//   * FFT   : as in frame_warp_kernel -- radix-E x radix-R2 packed real FFT in registers (FP32x2), one exchange through the warp's
//             private slice of shared memory, real split with the halving folded into the twiddle; window / twiddle tables in
//             shared memory.
//   * output: |X| or |X|^2 goes into a transposed CTA tile [B][TTP] (TTP = FW * odd >= TT: the tile writes of a warp -- G lanes
//             along the bins, FW frames -- fall on 32 distinct banks) and leaves as rows of TT consecutive frames per bin (the
//             reference layout is (1 + n_fft/2, T), frames innermost).  DB: two tiles, one CTA barrier per round and the drain of
//             round i overlaps the transforms of round i+1; otherwise one tile and two barriers.
//
// Zero padding (center=True, pad_mode='constant') is a predicate on the sample position; stale bytes of a stage are never used.
// Eligibility (launcher): constant padding, even hop, 16-byte aligned sample buffer, unit starts / valid lengths multiples of 4.
#pragma once

#include "syg_async.cuh"
#include "syg_frame_warp.cuh"

namespace sygdev {

template <class TL, int NW, bool DB>
struct RingGeom {
    using WT = WarpTile<TL, NW * 32>;
    static constexpr int M = TL::M, B = M + 1, FW = WT::FW, G = WT::G;
    static constexpr int NT = NW * 32;
    static constexpr int TT = NW * FW;                                  // frames per round
    static constexpr int TTP = FW * (NW | 1);                           // tile pitch: FW * odd >= TT
    // Z exchange region of one frame group.  64-bit shared accesses are served one half-warp at a time: the groups of a half-warp
    // (two for G = 8) must start 2G banks apart, so the pitch is a multiple of 32 words plus 2G mod 32 (WarpTile::RS, tuned for the
    // 32-bit |X|^2 stores of the feature kernel, is a multiple of 32 plus G: 2x the wavefronts on every exchange access here)
    static constexpr int ZW = 2 * WT::ZS;                                                        // words a group needs
    static constexpr int RSR = (FW == 1) ? (ZW + 3) / 4 * 4 : ((ZW - (2 * G) % 32 + 31) / 32 * 32 + (2 * G) % 32);
    static constexpr int WF = FW * RSR;
    static constexpr int kBarBytes = 64;
    static constexpr int kTableFloats = 2 * M + 2 * M + (M + 4);        // window, twiddles (transposed), half split twiddles
    static constexpr int kTileFloats = ((DB ? 2 : 1) * B * TTP + 3) / 4 * 4;
    SYG_HD static int stage_floats(int hop) { return ((TT - 1) * hop + 2 * M + 4 + 3) / 4 * 4; }
    static size_t bytes(int hop, int stages) {
        return (size_t)kBarBytes + sizeof(float) * ((size_t)kTableFloats + (size_t)NW * WF + kTileFloats + (size_t)stages * stage_floats(hop));
    }
};

template <class TL, int NW, bool DB, int S>
__global__ void __launch_bounds__(NW * 32, 1) stft_ring_kernel(const syg::FrameArgs a) {
    using RG = RingGeom<TL, NW, DB>;
    using WT = typename RG::WT;
    constexpr int E = WT::E, M = WT::M, G = WT::G, FW = WT::FW, R2 = WT::R2, LE = WT::LOG2E;
    constexpr int Q = E / R2, B = M + 1, NT = RG::NT, TT = RG::TT, TTP = RG::TTP, RSS = RG::RSR, WF = RG::WF;
    static_assert(R2 == G, "two-pass warp tile");
    constexpr bool kShflSplit = (SYG_SPLIT_SHFL != 0);                  // mirrors of the real split by SHFL (syg_device.cuh: mirror_of)
    constexpr bool kHalfZ = (SYG_SPLIT_HALF != 0);
    SYG_DYN_SMEM(smem_raw);
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem_raw);          // [S] "stage full"
    float* const fb = reinterpret_cast<float*>(smem_raw + RG::kBarBytes);
    float2* const t_win = reinterpret_cast<float2*>(fb);                                       // [M]     window pairs
    float2* const t_tw = t_win + M;                                                            // [R2][E] W_M^{r k}
    float2* const t_twsh = t_tw + M;                                                           // [M/2+1] 0.5 exp(-2 pi i k / n_fft)
    float* const regions = fb + RG::kTableFloats;
    float* const tile0 = regions + NW * WF;
    const int stage_floats = RG::stage_floats(a.hop);
    float* const stages = tile0 + RG::kTileFloats;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int f = lane / G, j = lane % G;
    float* const wbase = regions + warp * WF;
    float2* const zs = reinterpret_cast<float2*>(wbase + f * RSS);

    for (int i = tid; i < M; i += NT) t_win[i] = __ldg(reinterpret_cast<const float2*>(a.window) + i);
    // kHalfZ: pass 2 produces Z/2 (halved twiddles, element 0 entering with the pending scale 0.5), split in tangent form (split_power_h)
    for (int i = tid; i < M; i += NT) {
        const float2 w = __ldg(a.tw + (i / E) * (i % E));
        t_tw[i] = kHalfZ ? make_float2(0.5f * w.x, 0.5f * w.y) : w;
    }
    for (int i = tid; i <= M / 2; i += NT) {
        const float2 w = __ldg(a.tws + i);
        t_twsh[i] = kHalfZ ? split_twiddle_h(w, 4 * i < M) : make_float2(0.5f * w.x, 0.5f * w.y);
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncthreads();

    // rounds are numbered in 32 bits (the launcher sends larger problems to the other kernels); units are analytic (start = u * stride)
    const unsigned RU = (unsigned)((a.T + TT - 1) / TT);                                       // rounds per unit
    const unsigned n_rounds = (unsigned)a.g.n_units * RU;
    const unsigned stride = gridDim.x;
    auto unit_of = [&](unsigned u) {
        UnitRef r;
        r.start = (long long)u * a.g.unit_stride;
        long long v = a.g.total_len - r.start;
        v = v > a.g.unit_len ? a.g.unit_len : v;
        r.valid = v < 0 ? 0 : v;
        return r;
    };

    // producer side (thread 0): arm stage i % S and start the copy of local round i
    auto issue = [&](unsigned i) {
        const unsigned long long R64 = blockIdx.x + (unsigned long long)i * stride;
        if (R64 >= n_rounds) return;
        const unsigned R = (unsigned)R64;
        const unsigned u = R / RU;
        const int t0 = (int)(R - u * RU) * TT;
        const UnitRef ur = unit_of(u);
        const long long base = (long long)t0 * a.hop - a.cpad;
        const int nfr = min(TT, a.T - t0);
        long long lo = base > 0 ? base : 0;
        long long hi = base + (long long)(nfr - 1) * a.hop + 2 * M;
        hi = (hi + 3) & ~3LL;
        if (hi > ur.valid) hi = ur.valid;
        const int s = (int)(i % S);
        const unsigned bytes = hi > lo ? (unsigned)((hi - lo) * (long long)sizeof(float)) : 0u;
        mbar_expect_tx(&bars[s], bytes);
        if (bytes) bulk_g2s(stages + (size_t)s * stage_floats + (lo - base), a.y + ur.start + lo, bytes, &bars[s]);
    };
    if (tid == 0) {
        for (unsigned i = 0; i < (unsigned)S; ++i) issue(i);
    }

    for (unsigned i = 0;; ++i) {
        const unsigned long long R64 = blockIdx.x + (unsigned long long)i * stride;
        if (R64 >= n_rounds) break;                                                            // CTA uniform
        const unsigned R = (unsigned)R64;
        const unsigned u = R / RU;
        const int t0 = (int)(R - u * RU) * TT;
        const UnitRef ur = unit_of(u);
        const long long base = (long long)t0 * a.hop - a.cpad;
        const int nfr = min(TT, a.T - t0);
        const int s = (int)(i % S);
        float* const tl = tile0 + ((DB && (i & 1)) ? B * TTP : 0);
        const int fr = warp * FW + f;                                                          // this lane group's frame of the round
        const long long p0 = base + (long long)fr * a.hop;                                     // its first sample, in unit coordinates

        mbar_wait(&bars[s], (unsigned)((i / S) & 1));

        // ---------------- samples (shared memory) x window -> registers ----------------
        float2 z[E];
        {
            const float* src = stages + (size_t)s * stage_floats + fr * a.hop;                 // 8-byte aligned: hop is even
            const bool interior = fr < nfr && p0 >= 0 && p0 + 2 * M <= ur.valid;
            if (__all_sync(kFull, interior)) {
                const float2* s2 = reinterpret_cast<const float2*>(src);
                SYG_UNROLL
                for (int r = 0; r < E; ++r) {
                    const int c = j + r * G;
                    z[r] = __fmul2_rn(s2[c], t_win[c]);
                }
            } else {
                // frames that touch the padding (or lie past the unit's last frame): one unsigned compare per sample
                const long long lo64 = p0 < 0 ? -p0 : 0, hi64 = ur.valid - p0;
                const int lo = (int)(lo64 < 2 * M ? lo64 : 2 * M);
                const int hi = (fr < nfr) ? (int)(hi64 < 0 ? 0 : (hi64 < 2 * M ? hi64 : 2 * M)) : 0;
                const unsigned span = hi > lo ? (unsigned)(hi - lo) : 0u;
                SYG_UNROLL
                for (int r = 0; r < E; ++r) {
                    const int c = j + r * G;
                    const unsigned d = (unsigned)(2 * c - lo);
                    float2 v;
                    v.x = (d < span) ? src[2 * c] : 0.0f;
                    v.y = (d + 1u < span) ? src[2 * c + 1] : 0.0f;
                    z[r] = __fmul2_rn(v, t_win[c]);
                }
            }
        }

        // ---------------- pass 1: radix E; exchange; pass 2: twiddles + radix R2; natural order Z ----------------
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
                const float2 w = t_tw[r * E + k];
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

        // ---------------- real split -> |X| or |X|^2 -> transposed tile ----------------
        if (!DB) __syncthreads();                                                              // the single tile has been drained by everyone
        {
            const int jz = (j == 0) ? 1 : 0;
            const float2* const zk0 = zs + j;
            const float2* const zm0 = zs - j;
            const float2* const zm1 = zm0 + jz;
            float* const tcol = tl + fr;                                                       // column of this frame
            SYG_UNROLL
            for (int ii = 0; ii <= E / 2; ++ii) {
                const int kk = ii * G;
                const int k = j + kk;
                if (ii == E / 2 && j != 0) break;
                float2 zk, zm;
                if constexpr (kShflSplit) {                                                    // mirrors from the partner lane's registers
                    zk = z[zreg_of<E, G>(ii)];
                    zm = zk;                                                                   // ii = E/2 (lane 0): bin M/2 pairs with itself
                    if (ii < E / 2) zm = mirror_of<E, G>(z, ii, j);
                } else {
                    zk = zk0[kk + (kk >> LE)];
                    const int c1 = (M - kk) + ((M - kk - 1) >> LE);
                    const bool blk = ((M - kk) & (E - 1)) == 0;
                    zm = blk ? zm1[c1] : zm0[c1];
                    if (ii == 0 && j == 0) zm = zk;                                            // k = 0 pairs with itself (DC / Nyquist)
                }
                float pk_, pm_;
                if constexpr (kHalfZ) split_power_h(4 * ii < E, zk, zm, t_twsh[k], pk_, pm_);
                else split_power(zk, zm, t_twsh[k], pk_, pm_);
                if (a.out_kind == 1) { pk_ = sqrt_approx(pk_); pm_ = sqrt_approx(pm_); }
                tcol[k * TTP] = pk_;
                if (2 * k != M) tcol[(M - k) * TTP] = pm_;
            }
        }
        __syncthreads();                                                                       // tile complete; stage s consumed by every warp
        if (tid == 0) issue(i + S);

        // ---------------- drain: rows of nfr consecutive frames per bin ----------------
        {
            float* const obase = reinterpret_cast<float*>(a.stft_out) + (long long)u * B * a.T + t0;
            if ((a.T & 1) == 0 && (TT & 1) == 0 && (TTP & 1) == 0 && ((reinterpret_cast<uintptr_t>(obase) & 7u) == 0)) {
                // even T: every row starts 8-byte aligned -> one 64-bit store per frame pair (half the store instructions)
                constexpr int HP = TT / 2, KS2 = NT / HP;                                      // pairs per row, rows per pass
                const int sp = tid % HP, kq = tid / HP;
                if (2 * sp < nfr) {
                    const bool both = 2 * sp + 1 < nfr;
                    const float* src = tl + kq * TTP + 2 * sp;
                    float* dst = obase + (long long)kq * a.T + 2 * sp;
                    const long long dstep = (long long)KS2 * a.T;
                    SYG_UNROLL_BY(4)
                    for (int k = kq; k < B; k += KS2, src += KS2 * TTP, dst += dstep) {
                        const float2 v = *reinterpret_cast<const float2*>(src);
                        if (both) *reinterpret_cast<float2*>(dst) = v;
                        else *dst = v.x;
                    }
                }
            } else {
                constexpr int KS = NT / TT;                                                    // = 32 / FW rows per pass
                const int sl = tid % TT, kq = tid / TT;
                if (sl < nfr) {
                    const float* src = tl + kq * TTP + sl;
                    float* dst = obase + (long long)kq * a.T + sl;
                    const long long dstep = (long long)KS * a.T;
                    SYG_UNROLL_BY(4)
                    for (int k = kq; k < B; k += KS, src += KS * TTP, dst += dstep) *dst = *src;
                }
            }
        }
    }
}

}  // namespace sygdev
