// sygnals_b200/csrc/syg_stft_big.cuh
//
// stft_big_kernel<R, NW>: STFT magnitude / power output (compute_stft, sygnals/core/dsp.py:167-229) for n_fft = 2048 R, R = 2 or 4
// (n_fft 4096 / 8192), built from the warp-synchronous 1024-point register FFT of the smaller transforms.
//
// The packed frame z[n] = x[2n] + i x[2n+1], n < M = 1024 R, is decimated into R interleaved sub-sequences z_r[m] = z[R m + r].  Two
// warps own a frame (R of them: one sub-FFT each) (radix-32 x radix-32 in registers, FP32x2 butterflies, one exchange through a
// private shared-memory region, natural-order result Z_r left there), then -- after a CTA barrier -- the 64 lanes of the pair share
// the recombination: for a residue q < 1024
//     Z[q + 1024 a] = sum_r W_R^{r a} (W_M^{r q} Z_r[q])        (an R-point butterfly over the twiddled sub-spectra)
// and because M - (q + 1024 a) = (1024 - q) + 1024 (R-1-a), a lane that takes the residue pair (q, 1024 - q) holds both members of
// R real-split pairs in registers: X[k], X[M-k] -> |X| or |X|^2 go straight into the transposed CTA tile [B][TTP] and leave as rows
// of TT consecutive frames per bin.  Twiddles: W_M^q and W_{2M}^q come from the plan tables (read-only path), their powers by
// complex squaring, the factors of the mirrored residue and of the a-offsets are constants (powers of exp(-i pi / R)).
//
// The CTA-cooperative kernel this replaces (stft_tile_kernel: scalar three-pass FFT, a CTA barrier per pass) took 1.27 / 1.43 ms
// for the 4096-clip sweep point; see DESIGN.md for the measured numbers of this one.
#pragma once

#include "syg_frame_warp.cuh"

namespace sygdev {

template <int R, int NW>
struct BigGeom {
    using TL = FftTile<10, 32>;                                        // the 1024-point sub-transform
    using WT = WarpTile<TL, NW * 32>;
    static constexpr int MS = 1024;                                    // points of a sub-FFT
    static constexpr int M = MS * R, B = M + 1;
    static constexpr int NT = NW * 32;
    static constexpr int WPF = R;                                      // warps per frame: one sub-FFT each
    static constexpr int TT = NW / WPF;                                // frames per round
    static constexpr int TTP = TT | 1;                                 // odd tile pitch: bins along the lanes never collide
    // floats of one sub-spectrum region (zpad layout).  R = 4: regions 8 banks apart (ZF = 8 mod 32), so that the staging stores of a
    // half-warp -- points of sub-sequences r and r + 2 at the same index -- fall on disjoint banks
    static constexpr int ZF = (R == 4) ? 2120 : (2 * WT::ZS + 3) / 4 * 4;
    static_assert(ZF >= 2 * WT::ZS, "region holds the padded sub-spectrum");
    static constexpr int kTwFloats = 2 * MS;                           // transposed pass-2 twiddles of the sub-FFT
    static constexpr int kTileFloats = (B * TTP + 3) / 4 * 4;
    static constexpr size_t bytes = sizeof(float) * ((size_t)kTwFloats + (size_t)TT * R * ZF + kTileFloats);
    static_assert(R == 2 || R == 4, "n_fft 4096 or 8192");
};

// the warps that share a frame (named barrier 1 + slot); the CPU emulator has CTA barriers only, and every warp
// of the CTA reaches this point once per round, so a CTA barrier is an equivalent superset there
template <int THREADS>
SYG_DEVICE SYG_INLINE void frame_barrier(int slot) {
#ifdef SYG_EMU
    (void)slot;
    __syncthreads();
#else
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(THREADS) : "memory");
#endif
}

// (a + i b) * (c + i d)
SYG_DEVICE SYG_INLINE float2 cmulf(float2 x, float2 w) {
    return make_float2(__fmaf_rn(x.x, w.x, -x.y * w.y), __fmaf_rn(x.x, w.y, x.y * w.x));
}
SYG_DEVICE SYG_INLINE float2 cconj(float2 x) { return make_float2(x.x, -x.y); }
SYG_DEVICE SYG_INLINE float2 mul_mi(float2 x) { return make_float2(x.y, -x.x); }      // x * (-i)

// R-point DFT over r of t[r] (forward, W_R = exp(-2 pi i / R)), in place: t[a] = sum_r W_R^{r a} t[r]
template <int R>
SYG_DEVICE SYG_INLINE void dft_small(float2* t) {
    if (R == 2) {
        const float2 a = t[0], b = t[1];
        t[0] = __fadd2_rn(a, b);
        t[1] = __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
    } else {
        const float2 s02 = __fadd2_rn(t[0], t[2]), d02 = __ffma2_rn(t[2], make_float2(-1.0f, -1.0f), t[0]);
        const float2 s13 = __fadd2_rn(t[1], t[3]), d13 = mul_mi(__ffma2_rn(t[3], make_float2(-1.0f, -1.0f), t[1]));   // (t1 - t3) * (-i)
        t[0] = __fadd2_rn(s02, s13);
        t[2] = __ffma2_rn(s13, make_float2(-1.0f, -1.0f), s02);
        t[1] = __fadd2_rn(d02, d13);
        t[3] = __ffma2_rn(d13, make_float2(-1.0f, -1.0f), d02);
    }
}

template <int R, int NW>
__global__ void __launch_bounds__(NW * 32, 1) stft_big_kernel(const syg::FrameArgs a) {
    using BG = BigGeom<R, NW>;
    using WT = typename BG::WT;
    constexpr int E = 32, G = 32, MS = BG::MS, M = BG::M, B = BG::B, NT = BG::NT, TT = BG::TT, TTP = BG::TTP, WPF = BG::WPF, ZF = BG::ZF, LE = 5;
    constexpr int LPF = 32 * WPF;                                                              // lanes per frame
    SYG_DYN_SMEM(smem_raw);
    float* const fb = reinterpret_cast<float*>(smem_raw);
    float2* const t_tw = reinterpret_cast<float2*>(fb);                                        // [32][32] W_1024^{r k} (transposed)
    float* const regions = fb + BG::kTwFloats;                                                 // [TT][R][ZF]
    float* const tile = regions + (size_t)TT * R * ZF;                                         // [B][TTP]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int slot = warp / WPF, sub = warp % WPF;                                             // frame of the round, this warp's sub-sequence
    float* const freg = regions + (size_t)slot * R * ZF;                                       // this frame's R sub-spectra

    for (int i = tid; i < MS; i += NT) t_tw[i] = __ldg(a.tw1k + (i / E) * (i % E));
    __syncthreads();

    const long long n_rounds = (a.n_frames + TT - 1) / TT;
    for (long long round = blockIdx.x; round < n_rounds; round += gridDim.x) {
        const long long gf = round * TT + slot;
        const bool valid = gf < a.n_frames;
        const long long u = valid ? gf / a.T : 0;
        const int t = valid ? (int)(gf - u * a.T) : 0;
        UnitRef ur = unit_ref(a.g, u);
        if (!valid) ur.valid = 0;
        const long long p0 = (long long)t * a.hop - a.cpad;

        // ---------------- stage the windowed frame, de-interleaved, in the R regions ----------------
        // The two warps of the pair read the frame with coalesced 16-byte loads (a strided read of one sub-sequence would use a
        // quarter / half of every sector it touches) and scatter packed point n = R m + r to region r, index m (linear layout;
        // the sub-FFT's first pass then reads index lane + 32 i without conflicts and overwrites the region in the padded layout).
        {
            const int L = sub * 32 + lane;                                                     // lane within the frame's group of warps
            constexpr int C4 = M / (2 * LPF);                                                  // float4 chunks (two packed points) per lane
            const float* src = a.y + ur.start + p0;
            const bool interior = valid && p0 >= 0 && p0 + 2 * M <= ur.valid && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
            const float4* w4 = reinterpret_cast<const float4*>(a.window);
            auto put = [&](int n, float2 v) {                                                  // packed point n of the frame
                reinterpret_cast<float2*>(freg + (size_t)(n % R) * ZF)[n / R] = v;
            };
            if (__all_sync(kFull, interior)) {
                const float4* s4 = reinterpret_cast<const float4*>(src);
                SYG_UNROLL_BY(8)
                for (int i = 0; i < C4; ++i) {
                    const int c4 = L + LPF * i;
                    const float4 v = __ldg(s4 + c4), w = __ldg(w4 + c4);
                    put(2 * c4, make_float2(v.x * w.x, v.y * w.y));
                    put(2 * c4 + 1, make_float2(v.z * w.z, v.w * w.w));
                }
            } else {
                SYG_UNROLL_BY(4)
                for (int i = 0; i < C4; ++i) {
                    const int c4 = L + LPF * i;
                    const float4 w = __ldg(w4 + c4);
                    const float2 v0 = load_pair(a.y, ur, p0 + 4 * c4, a.pad_mode), v1 = load_pair(a.y, ur, p0 + 4 * c4 + 2, a.pad_mode);
                    put(2 * c4, make_float2(v0.x * w.x, v0.y * w.y));
                    put(2 * c4 + 1, make_float2(v1.x * w.z, v1.y * w.w));
                }
            }
        }
        frame_barrier<LPF>(slot);

        // ---------------- this warp's sub-FFT: z_r[m] = z[R m + r], r = sub ----------------
        {
            const int r = sub;
            float2* const zs = reinterpret_cast<float2*>(freg + (size_t)r * ZF);
            float2 z[E];
            SYG_UNROLL
            for (int i = 0; i < E; ++i) z[i] = zs[lane + i * G];
            __syncwarp();                                                                      // all inputs are in registers before the scatter
            dft_dif_p<E, 1>(z);
            SYG_UNROLL
            for (int kp = 0; kp < E; ++kp) zs[zpad<LE>(lane * E + kp)] = z[bitrev(kp, LE)];
            __syncwarp();
            SYG_UNROLL
            for (int rr = 0; rr < E; ++rr) z[rr] = zs[zpad<LE>(lane + rr * (MS / E))];
            __syncwarp();
            SYG_UNROLL
            for (int rr = 1; rr < E; ++rr) {
                const float2 w = t_tw[rr * E + lane];
                cmul(z[rr].x, z[rr].y, w.x, w.y);
            }
            dft_dif_p<E, 1>(z);
            SYG_UNROLL
            for (int kp = 0; kp < E; ++kp) zs[zpad<LE>(lane + kp * E)] = z[bitrev(kp, LE)];
        }
        __syncthreads();                                   // every sub-spectrum of the round is in place; the tile has been drained

        // ---------------- recombination + real split -> |X| or |X|^2 -> tile ----------------
        {
            const int L = sub * 32 + lane;                                                     // lane within the frame's group of warps
            const float2* Zr[R];
            SYG_UNROLL
            for (int r = 0; r < R; ++r) Zr[r] = reinterpret_cast<const float2*>(freg + (size_t)r * ZF);
            float* const tcol = tile + slot;
            const bool mag = (a.out_kind == 1);
            // constants: W_M^{1024} = exp(-2 pi i / R) (mirrored residue), exp(-i pi a / R) (real-split twiddle of the a-th copy)
            auto emit = [&](int k, float2 zk, float2 zm, float2 wh, bool both) {
                float pk_, pm_;
                split_power(zk, zm, wh, pk_, pm_);
                if (mag) { pk_ = sqrt_approx(pk_); pm_ = sqrt_approx(pm_); }
                tcol[k * TTP] = pk_;
                if (both) tcol[(M - k) * TTP] = pm_;
            };
            auto rot_a = [&](float2 w, int aidx) -> float2 {                                   // w * exp(-i pi aidx / R)
                if (aidx == 0) return w;
                if (R == 2) return mul_mi(w);                                                  // aidx = 1: -i
                if (aidx == 2) return mul_mi(w);
                const float c = 0.70710678118654752f;
                if (aidx == 1) return make_float2(c * (w.x + w.y), c * (w.y - w.x));           // * (1 - i)/sqrt2
                return make_float2(c * (w.y - w.x), -c * (w.x + w.y));                         // aidx = 3: * (-1 - i)/sqrt2
            };
            SYG_UNROLL
            for (int ii = 0; ii < 512 / LPF; ++ii) {
                const int q = L + LPF * ii;                                                    // 0..511
                if (q == 0) continue;                                                          // residues 0 and 512: below
                const int qm = MS - q;
                const float2 w1 = __ldg(a.tw + q);                                            // W_M^q
                const float2 wsh = __ldg(a.twsh + q);                                          // 0.5 exp(-2 pi i q / n_fft)
                float2 T[R], U[R];
                T[0] = Zr[0][zpad<LE>(q)];
                U[0] = Zr[0][zpad<LE>(qm)];
                float2 wp = w1;                                                                // W_M^{r q}
                SYG_UNROLL
                for (int r = 1; r < R; ++r) {
                    T[r] = cmulf(Zr[r][zpad<LE>(q)], wp);
                    // W_M^{r (1024 - q)} = (W_M^{1024})^r conj(W_M^{r q}),  W_M^{1024} = exp(-2 pi i / R)
                    float2 wc = cconj(wp);
                    if (R == 2) wc = make_float2(-wc.x, -wc.y);                                // r = 1: * (-1)
                    else if (r == 1) wc = mul_mi(wc);                                          // * (-i)
                    else if (r == 2) wc = make_float2(-wc.x, -wc.y);                           // * (-1)
                    else wc = make_float2(-wc.y, wc.x);                                        // r = 3: * (+i)
                    U[r] = cmulf(Zr[r][zpad<LE>(qm)], wc);
                    if (r + 1 < R) wp = cmulf(wp, w1);
                }
                dft_small<R>(T);                                                               // T[a] = Z[q + 1024 a]
                dft_small<R>(U);                                                               // U[a] = Z[(1024 - q) + 1024 a]
                SYG_UNROLL
                for (int aa = 0; aa < R; ++aa) emit(q + MS * aa, T[aa], U[R - 1 - aa], rot_a(wsh, aa), true);
            }
            if (L == 0) {
                // residue 0: Z[1024 a] from the untwiddled Z_r[0]; k = 0 pairs with itself (DC / Nyquist), k = 1024 a with 1024 (R - a)
                float2 T[R];
                SYG_UNROLL
                for (int r = 0; r < R; ++r) T[r] = Zr[r][0];
                dft_small<R>(T);
                const float2 w0 = __ldg(a.twsh);                                               // (0.5, 0)
                emit(0, T[0], T[0], w0, true);                                                 // X[0], X[M]
                if (R == 2) {
                    emit(MS, T[1], T[1], rot_a(w0, 1), false);                                 // k = M/2 pairs with itself
                } else {
                    emit(MS, T[1], T[3], rot_a(w0, 1), true);                                  // k = 1024 <-> 3072
                    emit(2 * MS, T[2], T[2], rot_a(w0, 2), false);                             // k = M/2
                }
                // residue 512: k = 512 + 1024 a pairs with 512 + 1024 (R-1-a)
                const float2 w1 = __ldg(a.tw + MS / 2);
                const float2 wsh = __ldg(a.twsh + MS / 2);
                float2 V[R];
                V[0] = Zr[0][zpad<LE>(MS / 2)];
                float2 wp = w1;
                SYG_UNROLL
                for (int r = 1; r < R; ++r) {
                    V[r] = cmulf(Zr[r][zpad<LE>(MS / 2)], wp);
                    if (r + 1 < R) wp = cmulf(wp, w1);
                }
                dft_small<R>(V);
                SYG_UNROLL
                for (int aa = 0; aa < R / 2; ++aa) emit(MS / 2 + MS * aa, V[aa], V[R - 1 - aa], rot_a(wsh, aa), true);
            }
        }
        __syncthreads();                                   // tile complete; the regions may be overwritten by the next round

        // ---------------- drain: rows of up to TT consecutive frames per bin ----------------
        // (one thread per row with 16-byte stores was measured: 0.73 vs 0.68 ms at n_fft 4096 -- a warp then writes 32 half sectors
        // per instruction instead of 4 full ones)
        {
            constexpr int KS = NT / TT;
            const int sl = tid % TT, kq = tid / TT;
            const long long gfd = round * TT + sl;
            if (gfd < a.n_frames) {
                const long long ud = gfd / a.T;
                const int td = (int)(gfd - ud * a.T);
                const float* src = tile + kq * TTP + sl;
                float* dst = reinterpret_cast<float*>(a.stft_out) + ((long long)ud * B + kq) * a.T + td;
                const long long dstep = (long long)KS * a.T;
                SYG_UNROLL_BY(4)
                for (int k = kq; k < B; k += KS, src += KS * TTP, dst += dstep) *dst = *src;
            }
        }
    }
}

}  // namespace sygdev
