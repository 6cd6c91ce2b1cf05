// sygnals_b200/csrc/syg_device.cuh
//
// Device building blocks shared by every kernel of the engine:
//   * in-register radix-2/4/8/16/32 DFT (decimation in frequency, compile-time twiddles),
//   * the shared-memory Stockham exchange used between register passes,
//   * CTA / thread-group scans and reductions (FP64 where the reference accumulates in float64),
//   * warp-level selection (bitonic sort, exact n-th order statistic) for spectral contrast.
//
// Tile model: a CTA of kThreads threads owns kThreads*E complex points per "round".  A frame of n_fft real
// samples is packed into M = n_fft/2 complex points (z[n] = x[2n] + i x[2n+1]); G = M/E threads cooperate on one
// frame and F = kThreads/G frames are transformed per round, all groups in lock step.
#pragma once

#include "syg_platform.h"
#include "syg_params.h"

namespace sygdev {

constexpr unsigned kFull = 0xffffffffu;

// loop unrolling by a compile-time factor (1 = keep the loop): the band-specialised kernels pass their trip counts as constants
#ifdef SYG_EMU
#define SYG_UNROLL_BY(n)
#else
#define SYG_PRAGMA_(x) _Pragma(#x)
#define SYG_UNROLL_BY(n) SYG_PRAGMA_(unroll n)
#endif

// --------------------------------------------------------------------------------------------------------
// shared-memory access wrappers (bank accounting in the emulator build; plain ld/st in the CUDA build)
// --------------------------------------------------------------------------------------------------------
#ifdef SYG_EMU
template <class T> SYG_INLINE T sld_(const T* p, int site) { ::sygemu::smem_access(p, sizeof(T), site); return *p; }
template <class T> SYG_INLINE void sst_(T* p, T v, int site) { ::sygemu::smem_access(p, sizeof(T), site); *p = v; }
#define SLD(p) ::sygdev::sld_((p), __LINE__)
#define SST(p, v) ::sygdev::sst_((p), (v), __LINE__)
#else
#define SLD(p) (*(p))
#define SST(p, v) (*(p) = (v))
#endif

// one padding word per 32: makes the stride-E writes of the first Stockham pass conflict free
SYG_DEVICE SYG_INLINE int padi(int i) { return i + (i >> 5); }
SYG_HD constexpr int padded_size(int n) { return n + (n >> 5) + 1; }

// --------------------------------------------------------------------------------------------------------
// compile-time twiddles  W_32^k = exp(-2 pi i k / 32)
// --------------------------------------------------------------------------------------------------------
SYG_DEVICE SYG_INLINE constexpr float cos32(int k) {
    constexpr float c[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                            0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f,
                            0.19509032201612825f, 0.0f};
    k &= 31;
    if (k > 16) k = 32 - k;           // cos is even
    return (k <= 8) ? c[k] : -c[16 - k];
}
SYG_DEVICE SYG_INLINE constexpr float sin32(int k) { return cos32(k - 8); }  // sin(x) = cos(x - pi/2)

// (xr, xi) *= W_R^K  with K, R compile-time after unrolling; trivial rotations cost no multiplies
template <int R>
SYG_DEVICE SYG_INLINE void mul_w(float& xr, float& xi, int k) {
    k &= (R - 1);
    const int k32 = k * (32 / R);
    if (k32 == 0) {
    } else if (k32 == 8) {            // * (-i)
        float t = xr; xr = xi; xi = -t;
    } else if (k32 == 16) {
        xr = -xr; xi = -xi;
    } else if (k32 == 24) {           // * (+i)
        float t = xr; xr = -xi; xi = t;
    } else if (k32 == 4) {            // (1 - i)/sqrt2
        float a = xr, b = xi; xr = (a + b) * 0.70710678118654752f; xi = (b - a) * 0.70710678118654752f;
    } else if (k32 == 12) {           // (-1 - i)/sqrt2
        float a = xr, b = xi; xr = (b - a) * 0.70710678118654752f; xi = -(a + b) * 0.70710678118654752f;
    } else if (k32 == 20) {           // (-1 + i)/sqrt2
        float a = xr, b = xi; xr = -(a + b) * 0.70710678118654752f; xi = (a - b) * 0.70710678118654752f;
    } else if (k32 == 28) {           // (1 + i)/sqrt2
        float a = xr, b = xi; xr = (a - b) * 0.70710678118654752f; xi = (a + b) * 0.70710678118654752f;
    } else {
        const float c = cos32(k32), s = -sin32(k32);   // W = c + i s
        float a = xr, b = xi;
        xr = __fmaf_rn(a, c, -b * s);
        xi = __fmaf_rn(a, s, b * c);
    }
}

SYG_HD constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
SYG_HD constexpr int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// In-register DFT of R points held at xr[o + i*s], xi[o + i*s] (i = 0..R-1): decimation in frequency, in place.
// Output X[k] ends up at position bitrev(k) (callers index with bitrev at compile time -> free).
template <int R, int S>
SYG_DEVICE SYG_INLINE void dft_dif(float* xr, float* xi) {
    SYG_UNROLL
    for (int half = R / 2; half >= 1; half >>= 1) {
        SYG_UNROLL
        for (int base = 0; base < R; base += 2 * half) {
            SYG_UNROLL
            for (int k = 0; k < half; ++k) {
                const int i0 = (base + k) * S, i1 = (base + k + half) * S;
                float ar = xr[i0], ai = xi[i0], br = xr[i1], bi = xi[i1];
                xr[i0] = ar + br; xi[i0] = ai + bi;
                float dr = ar - br, di = ai - bi;
                mul_w<R>(dr, di, k * (R / (2 * half)));
                xr[i1] = dr; xi[i1] = di;
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// packed variant: a complex point lives in one float2 (an aligned register pair), sums and differences of a butterfly
// are single FP32x2 instructions of sm_100 (add.f32x2 / fma.rn.f32x2 via __fadd2_rn / __ffma2_rn); rotations stay scalar
// on the two halves (a packed complex multiply would need the swapped pair).  Same DIF network as dft_dif.
// --------------------------------------------------------------------------------------------------------
template <int R>
SYG_DEVICE SYG_INLINE void rot_w(float2& d, int k) {            // d *= W_R^k
    mul_w<R>(d.x, d.y, k);
}

// Deferred twiddle scales (SYG_FFT_DEFER, default on).  A rotation by W = c + i s is written c (x - t y, t x + y) with t = s / c
// (or s (x ct - y, x + y ct) with ct = c / s where |s| > |c|, so |t| <= 1): two FFMA instead of FMUL + FFMA twice, and the factor c
// is NOT applied -- it stays a compile-time "pending scale" of that element.  The next butterfly absorbs it for free: a + rho b and
// a - rho b are single FFMA2 whose multiplier is a 32-bit immediate broadcast to both halves (FFMA2 R, R, imm, R), a factor common to
// both inputs simply carries on, a pending -1 is a sign flip of the constant.  In the radix-2 DIF network every pending scale is
// absorbed by the last stage (the element with twiddle exponent 0 of each block is never scaled); the tail loop below would apply a
// left-over one.  All bookkeeping (ps[], the ratios, the branches) is constant-folded after unrolling.  A radix-32 DFT takes
// 243 instructions instead of 307, a radix-16 one 91 instead of 111.
#ifndef SYG_FFT_DEFER
#define SYG_FFT_DEFER 1
#endif
SYG_DEVICE SYG_INLINE constexpr double cos32d(int k) {
    constexpr double c[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                             0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                             0.19509032201612826785, 0.0};
    k &= 31;
    if (k > 16) k = 32 - k;
    return (k <= 8) ? c[k] : -c[16 - k];
}
SYG_DEVICE SYG_INLINE constexpr double sin32d(int k) { return cos32d(k - 8); }
SYG_DEVICE SYG_INLINE constexpr bool scale_same(double a, double b) { return (a - b) < 1e-12 && (b - a) < 1e-12; }
SYG_DEVICE SYG_INLINE float2 bcast2(double v) { return make_float2((float)v, (float)v); }

// HALF0: element 0 enters with the pending scale 0.5 (SYG_FFT_DEFER only).  With every other input pre-multiplied by 0.5 (the
// halved pass-2 twiddle table) the network yields HALF the transform at no cost -- what split_power_h wants.
template <int R, int S, bool HALF0 = false>
SYG_DEVICE SYG_INLINE void dft_dif_p(float2* z) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
#if SYG_FFT_DEFER
    double ps[R];                                                   // pending scale of every element (compile time after unrolling)
    SYG_UNROLL
    for (int i = 0; i < R; ++i) ps[i] = 1.0;
    ps[0] = HALF0 ? 0.5 : 1.0;
    SYG_UNROLL
    for (int half = R / 2; half >= 1; half >>= 1) {
        SYG_UNROLL
        for (int base = 0; base < R; base += 2 * half) {
            SYG_UNROLL
            for (int k = 0; k < half; ++k) {
                const int i0 = (base + k) * S, i1 = (base + k + half) * S;
                const float2 a = z[i0], b = z[i1];
                const double pa = ps[base + k], pb = ps[base + k + half];
                const bool same = scale_same(pa, pb), a1 = scale_same(pa, 1.0), b1 = scale_same(pb, 1.0);
                const int kw = (k * (R / (2 * half))) & (R - 1);
                const int k32 = kw * (32 / R);
                // ---- sum: pa a + pb b
                if (same) { z[i0] = __fadd2_rn(a, b); ps[base + k] = pa; }
                else if (a1) { z[i0] = __ffma2_rn(b, bcast2(pb), a); ps[base + k] = 1.0; }
                else if (b1) { z[i0] = __ffma2_rn(a, bcast2(pa), b); ps[base + k] = 1.0; }
                else { z[i0] = __ffma2_rn(b, bcast2(pb / pa), a); ps[base + k] = pa; }
                // ---- difference (pa a - pb b) times W_R^kw
                double pd;
                if (k32 == 8 || k32 == 24) {                        // * (-i): (d.y, -d.x);  * (+i): (-d.y, d.x) -- the sign goes into the scale
                    const float rho = same ? 1.0f : (float)(pb / pa);
                    float2 d;
                    if (same) d = make_float2(a.y - b.y, b.x - a.x);
                    else d = make_float2(__fmaf_rn(-rho, b.y, a.y), __fmaf_rn(rho, b.x, -a.x));
                    z[i1] = d;
                    pd = (k32 == 8) ? pa : -pa;
                } else {
                    float2 d;
                    if (same) { d = __ffma2_rn(b, neg1, a); pd = pa; }
                    else if (a1) { d = __ffma2_rn(b, bcast2(-pb), a); pd = 1.0; }
                    else if (b1) { d = __ffma2_rn(a, bcast2(-pa), b); pd = -1.0; }           // b - pa a = -(pa a - b)
                    else { d = __ffma2_rn(b, bcast2(-pb / pa), a); pd = pa; }
                    if (k32 == 16) {
                        pd = -pd;
                    } else if (k32 != 0) {
                        const double c = cos32d(k32), sn = -sin32d(k32);                     // W = c + i sn
                        const double ac = c < 0.0 ? -c : c, as = sn < 0.0 ? -sn : sn;
                        const float x = d.x, y = d.y;
                        if (ac >= as) {                                                      // c ((x - t y) + i (t x + y))
                            const float t = (float)(sn / c);
                            d = make_float2(__fmaf_rn(-t, y, x), __fmaf_rn(t, x, y));
                            pd *= c;
                        } else {                                                             // sn ((ct x - y) + i (x + ct y))
                            const float ct = (float)(c / sn);
                            d = make_float2(__fmaf_rn(ct, x, -y), __fmaf_rn(ct, y, x));
                            pd *= sn;
                        }
                    }
                    z[i1] = d;
                }
                ps[base + k + half] = pd;
            }
        }
    }
    SYG_UNROLL
    for (int i = 0; i < R; ++i) {
        if (!scale_same(ps[i], 1.0)) z[i * S] = __fmul2_rn(z[i * S], bcast2(ps[i]));
    }
#else
    static_assert(!HALF0, "input scales need SYG_FFT_DEFER");
    SYG_UNROLL
    for (int half = R / 2; half >= 1; half >>= 1) {
        SYG_UNROLL
        for (int base = 0; base < R; base += 2 * half) {
            SYG_UNROLL
            for (int k = 0; k < half; ++k) {
                const int i0 = (base + k) * S, i1 = (base + k + half) * S;
                const float2 a = z[i0], b = z[i1];
                z[i0] = __fadd2_rn(a, b);
                const int kw = (k * (R / (2 * half))) & (R - 1);
                const int k32 = kw * (32 / R);
                if (k32 == 8) {                                   // (a - b) * (-i) = (d.y, -d.x): two scalar subtractions
                    z[i1] = make_float2(a.y - b.y, b.x - a.x);
                } else if (k32 == 16) {
                    z[i1] = __ffma2_rn(a, neg1, b);               // b - a
                } else if (k32 == 24) {                           // * (+i) = (-d.y, d.x)
                    z[i1] = make_float2(b.y - a.y, a.x - b.x);
                } else {
                    float2 d = __ffma2_rn(b, neg1, a);            // a - b
                    if (k32 != 0) rot_w<R>(d, kw);
                    z[i1] = d;
                }
            }
        }
    }
#endif
}

// complex multiply by a run-time twiddle (wr + i wi)
SYG_DEVICE SYG_INLINE void cmul(float& xr, float& xi, float wr, float wi) {
    float a = xr, b = xi;
    xr = __fmaf_rn(a, wr, -b * wi);
    xi = __fmaf_rn(a, wi, b * wr);
}

// --------------------------------------------------------------------------------------------------------
// FFT plan constants for a (LOG2M, E) tile
// --------------------------------------------------------------------------------------------------------
template <int LOG2M_, int E_>
struct FftTile {
    static constexpr int LOG2M = LOG2M_;
    static constexpr int E = E_;
    static constexpr int LOG2E = ilog2(E_);
    static constexpr int M = 1 << LOG2M_;                 // complex points per frame
    static constexpr int NFFT = 2 * M;                    // real samples per frame
    static constexpr int G = M / E_;                      // threads per frame
    static constexpr int F = kThreads / G;                // frames per round
    static constexpr int NPASS = (LOG2M_ + LOG2E - 1) / LOG2E;
    static constexpr int RLAST = 1 << (LOG2M_ - (NPASS - 1) * LOG2E);   // radix of the last pass
    static constexpr int MP = padded_size(M);             // padded complex stride of one frame in smem
    static constexpr int PW = padded_size(M + 1);         // padded power-spectrum stride of one frame
    static_assert(M >= E_, "frame too small for this tile");
    static_assert(G <= kThreads, "frame too large for this tile");
};

// One Stockham pass >= 2: gather E points from smem (stride M/R), twiddle, radix-R DFT, scatter back.
// tw[i] = exp(-2 pi i * i / M).  Two CTA barriers: all gathers complete before any scatter; scatters visible after.
template <class TL, int R, int NS>
SYG_DEVICE SYG_INLINE void stockham_pass(float* sre, float* sim, int fbase, int j, const float2* __restrict__ tw) {
    constexpr int E = TL::E, M = TL::M, G = TL::G, Q = E / R;
    float xr[E], xi[E];
    SYG_UNROLL
    for (int q = 0; q < Q; ++q) {
        const int b = j + q * G;
        SYG_UNROLL
        for (int r = 0; r < R; ++r) {
            const int idx = fbase + padi(b + r * (M / R));
            xr[q * R + r] = SLD(&sre[idx]);
            xi[q * R + r] = SLD(&sim[idx]);
        }
    }
    __syncthreads();
    SYG_UNROLL
    for (int q = 0; q < Q; ++q) {
        const int b = j + q * G;
        const int k = b & (NS - 1);
        constexpr int SH = TL::LOG2M - ilog2(NS * R);     // W_{NS*R}^{rk} = W_M^{rk << SH}
        SYG_UNROLL
        for (int r = 1; r < R; ++r) {
            const float2 w = __ldg(&tw[(r * k) << SH]);
            cmul(xr[q * R + r], xi[q * R + r], w.x, w.y);
        }
        dft_dif<R, 1>(xr + q * R, xi + q * R);
        const int base = fbase, ob = (b - k) * R + k;
        SYG_UNROLL
        for (int kp = 0; kp < R; ++kp) {
            const int src = q * R + bitrev(kp, ilog2(R));
            const int idx = base + padi(ob + kp * NS);
            SST(&sre[idx], xr[src]);
            SST(&sim[idx], xi[src]);
        }
    }
    __syncthreads();
}

// Full forward FFT of the CTA tile.  On entry xr/xi hold z[j + r*G] (r = 0..E-1) of frame f for each thread;
// on exit sre/sim[f*MP + padi(k)] hold Z[k] in natural order (after a CTA barrier).
template <class TL>
SYG_DEVICE SYG_INLINE void fft_tile_forward(float (&xr)[TL::E], float (&xi)[TL::E], float* sre, float* sim, int f, int j,
                                          const float2* __restrict__ tw) {
    constexpr int E = TL::E, LE = TL::LOG2E;
    const int fbase = f * TL::MP;
    // pass 1: radix E, NS = 1, no twiddles; thread j owns butterfly j and scatters to j*E + k'
    dft_dif<E, 1>(xr, xi);
    SYG_UNROLL
    for (int kp = 0; kp < E; ++kp) {
        const int idx = fbase + padi(j * E + kp);
        SST(&sre[idx], xr[bitrev(kp, LE)]);
        SST(&sim[idx], xi[bitrev(kp, LE)]);
    }
    __syncthreads();
    if constexpr (TL::NPASS >= 2) {
        if constexpr (TL::NPASS == 2) stockham_pass<TL, TL::RLAST, E>(sre, sim, fbase, j, tw);
        else stockham_pass<TL, E, E>(sre, sim, fbase, j, tw);
    }
    if constexpr (TL::NPASS >= 3) {
        if constexpr (TL::NPASS == 3) stockham_pass<TL, TL::RLAST, E * E>(sre, sim, fbase, j, tw);
        else stockham_pass<TL, E, E * E>(sre, sim, fbase, j, tw);
    }
    static_assert(TL::NPASS <= 3, "tile supports at most three passes");
}

// Real-input split for bin k (0 <= k <= M/2): given Z[k], Z[M-k] and Wk = exp(-2 pi i k / N) produce X[k], X[M-k].
SYG_DEVICE SYG_INLINE void real_split(float zkr, float zki, float zmr, float zmi, float wr, float wi,
                                    float& xkr, float& xki, float& xmr, float& xmi) {
    const float ar = 0.5f * (zkr + zmr), ai = 0.5f * (zki - zmi);     // A = (Zk + conj Zm)/2
    const float br = 0.5f * (zkr - zmr), bi = 0.5f * (zki + zmi);     // B = (Zk - conj Zm)/2
    const float cr = __fmaf_rn(br, wr, -bi * wi), ci = __fmaf_rn(br, wi, bi * wr);   // C = W^k B
    xkr = ar + ci; xki = ai - cr;                                     // X[k]   = A - i C
    xmr = ar - ci; xmi = -(ai + cr);                                  // X[M-k] = conj(A + i C)
}

// Power of the bin pair (k, M-k) straight from the packed pair (Z[k], Z[M-k]) and Wh = 0.5 exp(-2 pi i k / N):
//   S = Zk + Zm, D = Zk - Zm (one FP32x2 instruction each);  A = (S.x, D.y)/2, B = (D.x, S.y)/2, C = W B = Wh (D.x, S.y);
//   X[k] = A - iC = (S.x/2 + C.y, D.y/2 - C.x),  X[M-k] = conj(A + iC) = (S.x/2 - C.y, -(D.y/2 + C.x)).
SYG_DEVICE SYG_INLINE void split_power(float2 zk, float2 zm, float2 wh, float& pk, float& pm) {
    const float2 S = __fadd2_rn(zk, zm);
    const float2 D = __ffma2_rn(zm, make_float2(-1.0f, -1.0f), zk);
    const float cr = __fmaf_rn(D.x, wh.x, -S.y * wh.y), ci = __fmaf_rn(D.x, wh.y, S.y * wh.x);
    const float ur = __fmaf_rn(S.x, 0.5f, ci), ui = __fmaf_rn(D.y, 0.5f, -cr);
    const float vr = __fmaf_rn(S.x, 0.5f, -ci), vi = __fmaf_rn(D.y, 0.5f, cr);
    pk = __fmaf_rn(ur, ur, ui * ui);
    pm = __fmaf_rn(vr, vr, vi * vi);
}

// split_power on the HALVED packed spectrum (zk = Z[k]/2, zm = Z[M-k]/2, from dft_dif_p<.., HALF0> behind a halved twiddle table)
// with the split twiddle W = exp(-2 pi i k / N) = c + i s in tangent form: tk = (s/c, c) for k < M/4 (CF), (c/s, s) from M/4 on.
//   S = zk + zm, D = zk - zm;  A = (S.x, D.y), B = (D.x, S.y);  C = W B = kappa C',  C' = (B.x - t B.y, t B.x + B.y)  [CF]
//                                                                              or  C' = (t B.x - B.y, B.x + t B.y)
//   X[k] = A - iC = (S.x + kappa C'.y, D.y - kappa C'.x),  X[M-k] = conj(A + iC) = (S.x - kappa C'.y, -(D.y + kappa C'.x)).
// 12 instead of 14 instructions per bin pair (no halving multiplies, two FFMA for the rotation).
#ifndef SYG_SPLIT_HALF
#define SYG_SPLIT_HALF (SYG_FFT_DEFER)
#endif
SYG_HD inline float2 split_twiddle_h(float2 w, bool cform) {       // w = exp(-2 pi i k / N) as (cos, -sin)
    return cform ? make_float2(w.y / w.x, w.x) : make_float2(w.x / w.y, w.y);
}
SYG_DEVICE SYG_INLINE void split_power_h(bool CF, float2 zk, float2 zm, float2 tk, float& pk, float& pm) {   // CF: compile time after unrolling
    const float2 S = __fadd2_rn(zk, zm);
    const float2 D = __ffma2_rn(zm, make_float2(-1.0f, -1.0f), zk);
    float cx, cy;
    if (CF) { cx = __fmaf_rn(-tk.x, S.y, D.x); cy = __fmaf_rn(tk.x, D.x, S.y); }
    else { cx = __fmaf_rn(tk.x, D.x, -S.y); cy = __fmaf_rn(tk.x, S.y, D.x); }
    const float ur = __fmaf_rn(tk.y, cy, S.x), ui = __fmaf_rn(-tk.y, cx, D.y);
    const float vr = __fmaf_rn(-tk.y, cy, S.x), vi = __fmaf_rn(tk.y, cx, D.y);
    pk = __fmaf_rn(ur, ur, ui * ui);
    pm = __fmaf_rn(vr, vr, vi * vi);
}

// --------------------------------------------------------------------------------------------------------
// Mirror exchange of the real split by warp shuffles (two-pass warp tiles: R2 = G, Q = E / G).
// After pass 2 the lane j of a frame's G lanes holds Z[j + G m], m < E, in register zreg_of(m).  The split pairs bin
// k = j + G m with M - k = (G - j) + G (E - 1 - m): the mirror of a lane's lower half (m < E/2) is the UPPER half of lane
// G - j, register zreg_of(E - 1 - m) -- the same register in every lane, so one SHFL per word moves it.  Lane 0 is the one
// exception (M - G m = G (E - m): its own register zreg_of(E - m), m = 0 pairing with itself for DC / Nyquist): it selects
// that register before the shuffle and reads from itself.  Replaces the natural-order store of Z and the two strided loads of
// the split (E STS.64 + (E + 2) LDS.64 per lane: >= 4 E shared-memory wavefronts per warp task) by E SEL + E SHFL.
// --------------------------------------------------------------------------------------------------------
#ifndef SYG_SPLIT_SHFL
#define SYG_SPLIT_SHFL 1
#endif
template <int E, int G>
SYG_HD constexpr int zreg_of(int m) { return (m % (E / G)) * G + bitrev(m / (E / G), ilog2(G)); }

template <int E, int G>
SYG_DEVICE SYG_INLINE float2 mirror_of(const float2* z, int m, int j) {     // Z[M - (j + G m)], 0 <= m < E/2 (m: compile time)
    const float2 a = z[zreg_of<E, G>((E - m) % E)];
    const float2 b = z[zreg_of<E, G>(E - 1 - m)];
    const bool l0 = (j == 0);
    const float sx = l0 ? a.x : b.x, sy = l0 ? a.y : b.y;
    const int src = (G - j) & (G - 1);
    float2 r;
    r.x = __shfl_sync(kFull, sx, src, G);
    r.y = __shfl_sync(kFull, sy, src, G);
    return r;
}

// --------------------------------------------------------------------------------------------------------
// thread-group collectives.  Groups are G consecutive threads (G a power of two, 1..kThreads); every thread of
// the CTA calls them (lock step).  scratch: kThreads/32 doubles of shared memory.
// --------------------------------------------------------------------------------------------------------
template <int G>
SYG_DEVICE SYG_INLINE double group_sum(double v, double* scratch) {
    constexpr int W = G < 32 ? G : 32;
    SYG_UNROLL
    for (int o = W / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o, W);
    if constexpr (G > 32) {
        const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
        __syncthreads();
        if (lane == 0) scratch[warp] = v;
        __syncthreads();
        constexpr int WG = G / 32;
        const int w0 = (warp / WG) * WG;
        double t = 0.0;
        SYG_UNROLL
        for (int i = 0; i < WG; ++i) t += scratch[w0 + i];
        v = t;
    }
    return v;
}

template <int G>
SYG_DEVICE SYG_INLINE float group_max(float v, double* scratch) {
    constexpr int W = G < 32 ? G : 32;
    SYG_UNROLL
    for (int o = W / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o, W));
    if constexpr (G > 32) {
        const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
        float* fs = reinterpret_cast<float*>(scratch);
        __syncthreads();
        if (lane == 0) fs[warp] = v;
        __syncthreads();
        constexpr int WG = G / 32;
        const int w0 = (warp / WG) * WG;
        float t = fs[w0];
        SYG_UNROLL
        for (int i = 1; i < WG; ++i) t = fmaxf(t, fs[w0 + i]);
        v = t;
    }
    return v;
}

// inclusive scan over the group (thread order); returns this thread's inclusive prefix
template <int G>
SYG_DEVICE SYG_INLINE double group_scan_incl(double v, double* scratch) {
    constexpr int W = G < 32 ? G : 32;
    const int tid = threadIdx.x, lane = tid & 31, gl = lane & (W - 1);
    SYG_UNROLL
    for (int o = 1; o < W; o <<= 1) {
        double n = __shfl_up_sync(kFull, v, o, W);
        if (gl >= o) v += n;
    }
    if constexpr (G > 32) {
        const int warp = tid >> 5;
        __syncthreads();
        if (lane == 31) scratch[warp] = v;
        __syncthreads();
        constexpr int WG = G / 32;
        const int w0 = (warp / WG) * WG;
        double pre = 0.0;
        for (int i = w0; i < warp; ++i) pre += scratch[i];
        v += pre;
    }
    return v;
}

// --------------------------------------------------------------------------------------------------------
// warp-level selection
// --------------------------------------------------------------------------------------------------------
// full-warp bitonic sort, descending by lane (lane 0 gets the largest)
SYG_DEVICE SYG_INLINE float warp_sort_desc(float v) {
    const int lane = threadIdx.x & 31;
    SYG_UNROLL
    for (int k = 2; k <= 32; k <<= 1) {
        SYG_UNROLL
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            const float o = __shfl_xor_sync(kFull, v, jj);
            const bool desc = ((lane & k) == 0);          // block direction (k == 32: descending everywhere)
            const bool lower = ((lane & jj) == 0);
            v = (lower == desc) ? fmaxf(v, o) : fminf(v, o);
        }
    }
    return v;
}

SYG_DEVICE SYG_INLINE float warp_sum(float v) {
    SYG_UNROLL
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
SYG_DEVICE SYG_INLINE float warp_max(float v) {
    SYG_UNROLL
    for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
SYG_DEVICE SYG_INLINE int warp_excl_scan(int v, int& total) {
    const int lane = threadIdx.x & 31;
    int x = v;
    SYG_UNROLL
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += n;
    }
    total = __shfl_sync(kFull, x, 31);
    return x - v;
}

// Mean of sqrt() of the n largest (SIGN=+1) or n smallest (SIGN=-1) of the `count` non-negative values
// p[padi(lo + i)], i < count, taken by one full warp.  Selection is exact (ties irrelevant: only values
// enter the mean).  cand: 32 floats of warp-private shared memory.
template <int SIGN>
SYG_DEVICE SYG_INLINE float warp_extreme_mean_sqrt(const float* p, int lo, int count, int n, float* cand) {
    const int lane = threadIdx.x & 31;
    if (count <= 0) return __uint_as_float(0x7fc00000u);   // mean of nothing -> NaN (numpy)
    if (n > count) n = count;
    const float NEG = -3.0e38f;
    // key(x) = SIGN * x : "largest key" selection in both directions
    float lmax = NEG;
    int have = 0;
    for (int i = lane; i < count; i += 32) {
        const float key = SIGN * SLD(&p[padi(lo + i)]);
        lmax = fmaxf(lmax, key);
        have = 1;
    }
    (void)have;
    float thr = NEG;
    bool fast = (n <= 32);
    if (fast && count > 32) {
        const float srt = warp_sort_desc(lmax);
        thr = __shfl_sync(kFull, srt, n - 1);              // >= n values are >= thr
    }
    int c_total = count;
    if (fast) {
        int cnt = 0;
        if (count > 32) {
            for (int i = lane; i < count; i += 32) cnt += (SIGN * SLD(&p[padi(lo + i)]) >= thr) ? 1 : 0;
        } else {
            cnt = (lane < count) ? 1 : 0;
        }
        const int base = warp_excl_scan(cnt, c_total);
        if (c_total <= 32) {
            int slot = base;
            for (int i = lane; i < count; i += 32) {
                const float key = SIGN * SLD(&p[padi(lo + i)]);
                if (key >= thr) cand[slot++] = key;
            }
            __syncwarp();
            float v = (lane < c_total) ? cand[lane] : NEG;
            __syncwarp();
            v = warp_sort_desc(v);
            const float contrib = (lane < n) ? sqrtf(SIGN * v) : 0.0f;
            return warp_sum(contrib) / (float)n;
        }
    }
    // exact slow path (n > 32 or many ties): bitwise search for the n-th largest key, on an order-preserving
    // unsigned image of the float key
    auto okey = [](float k) -> unsigned {
        unsigned u = __float_as_uint(k);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    };
    unsigned prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned trial = prefix | (1u << bit);
        int cnt = 0;
        for (int i = lane; i < count; i += 32) cnt += (okey(SIGN * SLD(&p[padi(lo + i)])) >= trial) ? 1 : 0;
        cnt = __reduce_add_sync(kFull, cnt);
        if (cnt >= n) prefix = trial;
    }
    // prefix == okey(n-th largest key)
    float s_gt = 0.0f;
    int c_gt = 0;
    float tval = 0.0f;
    for (int i = lane; i < count; i += 32) {
        const float key = SIGN * SLD(&p[padi(lo + i)]);
        const unsigned ok = okey(key);
        if (ok > prefix) { s_gt += sqrtf(SIGN * key); c_gt++; }
        if (ok == prefix) tval = SIGN * key;
    }
    s_gt = warp_sum(s_gt);
    c_gt = __reduce_add_sync(kFull, c_gt);
    tval = warp_max(tval);                                  // all lanes holding it agree; others contribute 0 (values >= 0)
    return (s_gt + (float)(n - c_gt) * sqrtf(tval)) / (float)n;
}

// --------------------------------------------------------------------------------------------------------
// fire-and-forget maximum on a global word.  atomicMax() called by a single lane still goes through the compiler's warp
// aggregation (VOTE + REDUX + leader election: 7 instructions for one active lane); `red` is the one instruction that is needed.
// --------------------------------------------------------------------------------------------------------
SYG_DEVICE SYG_INLINE void red_max_u32(unsigned* p, unsigned v) {
#if defined(SYG_EMU)
    atomicMax(p, v);
#else
    asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// --------------------------------------------------------------------------------------------------------
// hardware square root (flush-to-zero: a denormal |X|^2 is silence)
// --------------------------------------------------------------------------------------------------------
SYG_DEVICE SYG_INLINE float sqrt_approx(float x) {
#if defined(SYG_EMU)
    return std::sqrt(x);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// --------------------------------------------------------------------------------------------------------
// warp kernel: spectral-contrast selection (everything in registers).  P spectra of the warp kernel use 4 pad words per 32 bins:
// --------------------------------------------------------------------------------------------------------
SYG_DEVICE SYG_INLINE int ppad(int k) { return k + ((k >> 5) << 2); }

// peak / valley of one spectral-contrast band for a full warp: mean magnitude of the n largest / n smallest |X|^2 bins
// of p[ppad(lo + i)], i < count, 1 <= n <= count.
//
// Streaming selection.  The band is lane-striped (element i belongs to lane i % 32, consecutive elements of a lane are
// 36 words apart in the ppad layout).  While streaming the band once from shared memory every lane keeps the 4 largest
// keys of BOTH directions sorted in registers (top: key = bits(x); bottom: key = ~bits(x); order preserving for
// non-negative floats; 0 = "nothing").  Extraction then works on the lanes' sorted heads:
//   bulk round : M2 = warp max of the second entries; every head > M2 is larger than all non-head elements, so the
//                set {a0 > M2} is exactly the |S| largest remaining elements -> all of them pop in one step
//                (each lane sums its own popped values; one warp reduction at the end);
//   single pop : when a bulk round would overshoot n or is blocked by a tie, the warp maximum pops alone; ties are
//                counted by value, so silent frames finish in one step.
// A lane that has popped its 4 tracked keys re-streams its elements below the last popped key (rare).  Exact.
SYG_DEVICE SYG_INLINE void cex(unsigned& lo_, unsigned& hi_) {        // compare-exchange: lo_ <= hi_ afterwards
    const unsigned l = min(lo_, hi_), h = max(lo_, hi_);
    lo_ = l; hi_ = h;
}
// insert v into the descending list t0 >= t1 >= t2 >= t3 (largest four)
SYG_DEVICE SYG_INLINE void ins4_desc(unsigned& t0, unsigned& t1, unsigned& t2, unsigned& t3, unsigned v) {
    unsigned h;
    h = max(t0, v); v = min(t0, v); t0 = h;
    h = max(t1, v); v = min(t1, v); t1 = h;
    h = max(t2, v); v = min(t2, v); t2 = h;
    t3 = max(t3, v);
}
// insert v into the ascending list t0 <= t1 <= t2 <= t3 (smallest four)
SYG_DEVICE SYG_INLINE void ins4_asc(unsigned& t0, unsigned& t1, unsigned& t2, unsigned& t3, unsigned v) {
    unsigned l;
    l = min(t0, v); v = max(t0, v); t0 = l;
    l = min(t1, v); v = max(t1, v); t1 = l;
    l = min(t2, v); v = max(t2, v); t2 = l;
    t3 = min(t3, v);
}

struct Sel4 {                        // one lane's view of a band: four largest (a, descending) and four smallest (b, ascending)
    unsigned a0, a1, a2, a3, b0, b1, b2, b3;
};

// stream this lane's `mine` elements (q[36 i]) once, keeping both sorted quadruples.  Elements are taken four at a time:
// a 5-exchange sorting network orders the quad, two bitonic half-merges fold it into the lists (8.5 min/max per element
// instead of 14 for element-wise insertion).
template <bool ST>
SYG_DEVICE SYG_INLINE Sel4 track4(const float* __restrict__ q, int mine, int count) {
#ifndef SYG_TRACK_UNROLL
#define SYG_TRACK_UNROLL 8
#endif
    constexpr int UQ = ST ? SYG_TRACK_UNROLL : 1, US = ST ? 4 : 1;             // static band layout: trip counts are constants -> straight-line code
    Sel4 t;
    t.a0 = t.a1 = t.a2 = t.a3 = 0u;
    t.b0 = t.b1 = t.b2 = t.b3 = 0xffffffffu;
    // warp-uniform trip counts: every lane owns at least count/32 elements and at most one more
    const int nq = (count >> 5) >> 2;                           // full quads every lane has
    const int nmax = (count + 31) >> 5;
    SYG_UNROLL_BY(UQ)
    for (int i = 0; i < nq; ++i) {
        const float* e = q + 144 * i;
        unsigned s0 = __float_as_uint(e[0]), s1 = __float_as_uint(e[36]), s2 = __float_as_uint(e[72]), s3 = __float_as_uint(e[108]);
        cex(s0, s1); cex(s2, s3); cex(s0, s2); cex(s1, s3); cex(s1, s2);          // s0 <= s1 <= s2 <= s3
        // largest four of {a} U {s}: max(a_i, s_i) is bitonic and holds them; two exchange stages sort it descending
        unsigned l0 = max(t.a0, s0), l1 = max(t.a1, s1), l2 = max(t.a2, s2), l3 = max(t.a3, s3);
        cex(l2, l0); cex(l3, l1); cex(l1, l0); cex(l3, l2);
        t.a0 = l0; t.a1 = l1; t.a2 = l2; t.a3 = l3;
        // smallest four of {b} U {s}: min(b_i, s_{3-i}); sort ascending
        unsigned m0 = min(t.b0, s3), m1 = min(t.b1, s2), m2 = min(t.b2, s1), m3 = min(t.b3, s0);
        cex(m0, m2); cex(m1, m3); cex(m0, m1); cex(m2, m3);
        t.b0 = m0; t.b1 = m1; t.b2 = m2; t.b3 = m3;
    }
    SYG_UNROLL_BY(US)
    for (int i = 4 * nq; i < nmax; ++i) {                        // <= 4 steps; the last one may be missing in some lanes
        const bool have = i < mine;
        const unsigned x = have ? __float_as_uint(q[36 * i]) : 0u;
        ins4_desc(t.a0, t.a1, t.a2, t.a3, x);
        ins4_asc(t.b0, t.b1, t.b2, t.b3, have ? x : 0xffffffffu);
    }
    return t;
}

// rebuild a lane's descending quadruple from its elements below `last` (keys = bits ^ flip) plus the copies of `last`
// it has not popped yet; popped = number of elements this lane has popped so far (they are its `popped` largest keys)
SYG_DEVICE SYG_INLINE uint4 refill4(const float* __restrict__ q, int mine, unsigned flip, unsigned last, int popped, unsigned tag) {
    unsigned n0 = 0u, n1 = 0u, n2 = 0u, n3 = 0u;
    int ge = 0;
#ifndef SYG_EMU
#pragma unroll 1
#endif
    for (int i = 0; i < mine; ++i) {
        const unsigned x = ((__float_as_uint(q[36 * i]) ^ flip) & ~31u) | tag;
        ge += (x >= last) ? 1 : 0;
        ins4_desc(n0, n1, n2, n3, (x < last) ? x : 0u);
    }
    for (int r = min(ge - popped, 4); r > 0; --r) ins4_desc(n0, n1, n2, n3, last);
    return make_uint4(n0, n1, n2, n3);
}

// returns {peak, valley}: mean of sqrt over the n largest / n smallest values of the band.  ST: (lo, count, n) are compile-time
// constants at the call site (band-specialised kernel): every loop below unrolls completely.
template <bool ST>
SYG_DEVICE SYG_INLINE float2 band_peak_valley_stream(const float* __restrict__ p, int lo, int count, int n) {
#ifndef SYG_POP_UNROLL
#define SYG_POP_UNROLL 3     // measured on B200 (cfg4): 3 / 5 / 8 / 32 -> 6.540 / 6.571 / 6.586 / 6.584 ms per 2 h (instruction-cache footprint)
#endif
    constexpr int UP = ST ? SYG_POP_UNROLL : 1;
    const int lane = threadIdx.x & 31;
    const float* q = p + ppad(lo + lane);
    const int mine = max((count - lane + 31) >> 5, 0);         // elements this lane owns
    float2 r;
    if (n == 1) {                                               // extremes of the band
        unsigned mx = 0u, mn = 0xffffffffu;
        if (count <= 64) {                                      // the usual case (low octave bands): two predicated loads, no loop
            const bool h0 = lane < count, h1 = lane + 32 < count;
            const unsigned x0 = h0 ? __float_as_uint(q[0]) : 0u, x1 = h1 ? __float_as_uint(q[36]) : 0u;
            mx = max(x0, x1);
            mn = min(h0 ? x0 : 0xffffffffu, h1 ? x1 : 0xffffffffu);
        } else {
            const int nmax = (count + 31) >> 5;                 // warp-uniform trip count; the last element may be missing in some lanes
            SYG_UNROLL_BY(UP)
            for (int i = 0; i < nmax; ++i) {
                const bool have = i < mine;
                const unsigned x = have ? __float_as_uint(q[36 * i]) : 0u;
                mx = max(mx, x);
                mn = min(mn, have ? x : 0xffffffffu);
            }
        }
        r.x = sqrt_approx(__uint_as_float(__reduce_max_sync(kFull, mx)));
        r.y = sqrt_approx(__uint_as_float(__reduce_min_sync(kFull, mn)));
        return r;
    }
    Sel4 t;
    if (count <= 128) {
        // every lane owns at most four elements: two 5-exchange sorts give both sorted quadruples directly (missing elements are
        // 0 for the largest-first list and all-ones for the smallest-first list), no insertion loop
        const bool h0 = lane < count, h1 = lane + 32 < count, h2 = lane + 64 < count, h3 = lane + 96 < count;
        const unsigned e0 = h0 ? __float_as_uint(q[0]) : 0u, e1 = h1 ? __float_as_uint(q[36]) : 0u;
        const unsigned e2 = h2 ? __float_as_uint(q[72]) : 0u, e3 = h3 ? __float_as_uint(q[108]) : 0u;
        unsigned s0 = e0, s1 = e1, s2 = e2, s3 = e3;
        cex(s0, s1); cex(s2, s3); cex(s0, s2); cex(s1, s3); cex(s1, s2);          // ascending
        t.a0 = s3; t.a1 = s2; t.a2 = s1; t.a3 = s0;
        unsigned r0 = h0 ? e0 : 0xffffffffu, r1 = h1 ? e1 : 0xffffffffu, r2 = h2 ? e2 : 0xffffffffu, r3 = h3 ? e3 : 0xffffffffu;
        cex(r0, r1); cex(r2, r3); cex(r0, r2); cex(r1, r3); cex(r1, r2);
        t.b0 = r0; t.b1 = r1; t.b2 = r2; t.b3 = r3;
    } else {
        t = track4<ST>(q, mine, count);
    }
    // both directions as "largest key first": top keys = bits, bottom keys = ~bits (0 = nothing).  The low 5 bits of every key
    // are replaced by the lane number: keys of different lanes never tie, so one REDUX names the single lane that pops and the
    // pop needs no ballot / lowest-lane election (17 -> 8 instructions per popped element and direction).  The price is a
    // selection and a value that are exact only down to 2^-18 relative (|X|^2; 2^-19 on the magnitude = 1.7e-5 dB, the parity
    // bar is 1e-3 dB and the FP32 spectrum itself carries ~1e-6).
    const unsigned tag = (unsigned)lane;
    unsigned a0 = (t.a0 & ~31u) | tag, a1 = (t.a1 & ~31u) | tag, a2 = (t.a2 & ~31u) | tag, a3 = (t.a3 & ~31u) | tag;
    unsigned b0 = (~t.b0 & ~31u) | tag, b1 = (~t.b1 & ~31u) | tag, b2 = (~t.b2 & ~31u) | tag, b3 = (~t.b3 & ~31u) | tag;
    int ra = n, rb = n;                                         // elements still to pop (warp uniform)
    float sa = 0.0f, sb = 0.0f;
    // Untracked elements are <= their lane's 4th key <= T = max over lanes of the 4th keys, so the tracked keys above T
    // are exactly the largest elements of the band.  If there are at least n of them (the common case), the n pops never
    // leave the tracked quadruples and the lean loop needs no bookkeeping; otherwise (strongly clustered spectra) the
    // full loop counts pops per lane and rebuilds the quadruple of a lane that runs dry.
    // (a lane can contribute at most n elements to the n extremes: with n <= 4 the pops stay inside its quadruple by construction)
    bool lean = ((count + 31) >> 5) <= 4 || n <= 4;
    if (!lean) {
        const unsigned ta = __reduce_max_sync(kFull, a3), tb = __reduce_max_sync(kFull, b3);
        const int ca = (a0 > ta ? 1 : 0) + (a1 > ta ? 1 : 0) + (a2 > ta ? 1 : 0);
        const int cb = (b0 > tb ? 1 : 0) + (b1 > tb ? 1 : 0) + (b2 > tb ? 1 : 0);
        const unsigned cc = __reduce_add_sync(kFull, (unsigned)(ca | (cb << 16)));
        lean = (int)(cc & 0xffffu) >= n && (int)(cc >> 16) >= n;
    }
    if (lean) {
        // one element per step and direction, both directions in one straight-line body: two independent REDUX chains for the
        // scheduler to interleave.  Pops are selects, not branches.
        // (bulk rounds -- pop every head above the warp maximum of the second keys at once, ~7 per round -- were measured in
        // round 1: two more REDUX per round cost more than the saved rounds, +1.3 % kernel time.  Rejected.)
        // The 4th key needs a refill (the bare lane tag: smaller than every real key) only where a lane may pop MORE than its four
        // tracked keys' worth: n > 4 pops from lanes that own at most four elements.  With n <= 4 at most four pops come from one
        // lane (the stale copies of the 4th key that the shifts leave behind are never reached), and under the T criterion above
        // every popped key is > T >= each lane's 4th key, so a 4th key never pops at all.  Saves two selects per popped pair.
        const bool refill = n > 4 && ((count + 31) >> 5) <= 4;
        SYG_UNROLL_BY(UP)
        for (int it = 0; it < n; ++it) {
            const unsigned ga = __reduce_max_sync(kFull, a0);
            const unsigned gb = __reduce_max_sync(kFull, b0);
            const bool oa = (a0 == ga), ob = (b0 == gb);
            sa += sqrt_approx(__uint_as_float(ga));
            sb += sqrt_approx(__uint_as_float(~gb));
            a0 = oa ? a1 : a0; a1 = oa ? a2 : a1; a2 = oa ? a3 : a2;
            b0 = ob ? b1 : b0; b1 = ob ? b2 : b1; b2 = ob ? b3 : b2;
            if (refill) { a3 = oa ? tag : a3; b3 = ob ? tag : b3; }
        }
    } else {
        int ha = min(mine, 4), hb = ha;                         // tracked keys left
        int pa = 0, pb = 0;                                     // elements popped by this lane
#ifndef SYG_EMU
#pragma unroll 1
#endif
        while (ra > 0 || rb > 0) {
            if (ra > 0) {
                const unsigned g = __reduce_max_sync(kFull, a0);
                const bool own = (a0 == g);
                const int c = min(__popc(__ballot_sync(kFull, own)), ra);
                sa = __fmaf_rn((float)c, sqrt_approx(__uint_as_float(g)), sa);
                ra -= c;
                a0 = own ? a1 : a0; a1 = own ? a2 : a1; a2 = own ? a3 : a2; a3 = own ? 0u : a3;
                pa += own ? 1 : 0;
                ha -= own ? 1 : 0;
                if (__any_sync(kFull, ha == 0 && pa < mine) && ra > 0) {
                    if (ha == 0 && pa < mine) {
                        const uint4 v = refill4(q, mine, 0u, g, pa, tag);
                        a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w;
                        ha = min(mine - pa, 4);
                    }
                }
            }
            if (rb > 0) {
                const unsigned g = __reduce_max_sync(kFull, b0);
                const bool own = (b0 == g);
                const int c = min(__popc(__ballot_sync(kFull, own)), rb);
                sb = __fmaf_rn((float)c, sqrt_approx(__uint_as_float(~g)), sb);
                rb -= c;
                b0 = own ? b1 : b0; b1 = own ? b2 : b1; b2 = own ? b3 : b2; b3 = own ? 0u : b3;
                pb += own ? 1 : 0;
                hb -= own ? 1 : 0;
                if (__any_sync(kFull, hb == 0 && pb < mine) && rb > 0) {
                    if (hb == 0 && pb < mine) {
                        const uint4 v = refill4(q, mine, 0xffffffffu, g, pb, tag);
                        b0 = v.x; b1 = v.y; b2 = v.z; b3 = v.w;
                        hb = min(mine - pb, 4);
                    }
                }
            }
        }
    }
    r.x = __fdividef(sa, (float)n);                             // MUFU.RCP + FMUL (2 ulp; the parity bar is 1e-3 dB)
    r.y = __fdividef(sb, (float)n);
    return r;
}

SYG_DEVICE SYG_INLINE void band_peak_valley_any(const float* __restrict__ p, int lo, int count, int n, float& peak, float& valley) {
    if (count <= 0) { peak = valley = __uint_as_float(0x7fc00000u); return; }     // mean of nothing -> NaN (numpy)
    if (n > count) n = count;
    if (n < 1) n = 1;
    const float2 r = band_peak_valley_stream<false>(p, lo, count, n);
    peak = r.x;
    valley = r.y;
}

}  // namespace sygdev
