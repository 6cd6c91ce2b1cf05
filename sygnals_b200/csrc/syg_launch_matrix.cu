// Matrix-level entry points of the boundary: the reference's functions that take a spectrogram instead of samples.
//   mfcc(S=log-mel)            sygnals/core/features/cepstral.py:94-117 -> librosa.feature.mfcc(S=...) = scipy.fftpack.dct(S, axis=-2,
//                              type, norm)[:n_mfcc] (+ sinusoidal lifter)
//   spectral_contrast(S=|X|)   sygnals/core/features/frequency_domain.py:147-212 -> librosa.feature.spectral_contrast(S=...)
// Both matrices are frequency-major (rows = mel bands / bins, columns = frames), as the reference holds them.
#include "syg_launch_common.h"
#include "syg_device.cuh"

namespace sygdev {

// out[u][c][t] = sum_n D[c][n] S[u][n][t]: one thread per output, coalesced along t, FP64 like scipy
__global__ void __launch_bounds__(kThreads) dct_matrix_kernel(const double* __restrict__ S, const double* __restrict__ D, long long n_units, int N,
                                                              long long T, int C, double* __restrict__ out) {
    const long long total = n_units * C * T;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += stride) {
        const long long t = idx % T;
        const long long uc = idx / T;
        const int c = (int)(uc % C);
        const long long u = uc / C;
        const double* s = S + u * N * T + t;
        const double* d = D + (long long)c * N;
        double acc = 0.0;
        for (int n = 0; n < N; ++n) acc = fma(__ldg(d + n), __ldg(s + (long long)n * T), acc);
        out[idx] = acc;
    }
}

struct ContrastSpecArgs {
    const float* S;              // [B][T] magnitudes
    int B;
    long long T;
    int nb;
    int band_lo[syg::kMaxBands], band_cnt[syg::kMaxBands], band_n[syg::kMaxBands];
    float* cws;                  // [T][2 nb] linear peaks | valleys
    unsigned* unit_max;          // [4]: slot 1 max peak, slot 2 max valley (bit images)
    int pw;                      // floats of one warp's padded spectrum
};

// one warp per frame: column t of S -> |X|^2 in the padded layout of the feature kernel -> the same selection code
__global__ void __launch_bounds__(kThreads) contrast_spectrum_kernel(const ContrastSpecArgs a) {
    SYG_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* const P = reinterpret_cast<float*>(smem_raw) + (size_t)warp * a.pw;
    const long long n_warps = (long long)gridDim.x * (kThreads / 32);
    float pmx = 0.0f, vmx = 0.0f;
    for (long long t = (long long)blockIdx.x * (kThreads / 32) + warp; t < a.T; t += n_warps) {
        for (int k = lane; k < a.B; k += 32) {
            const float m = __ldg(a.S + (long long)k * a.T + t);
            P[ppad(k)] = m * m;
        }
        __syncwarp();
        float mine = 0.0f;
        for (int bd = 0; bd < a.nb; ++bd) {
            float peak, valley;
            band_peak_valley_any(P, a.band_lo[bd], a.band_cnt[bd], a.band_n[bd], peak, valley);
            mine = (lane == bd) ? peak : ((lane == a.nb + bd) ? valley : mine);
            pmx = fmaxf(pmx, peak);
            vmx = fmaxf(vmx, valley);
        }
        if (lane < 2 * a.nb) a.cws[t * (2 * a.nb) + lane] = mine;
        __syncwarp();
    }
    if (lane == 0) {
        if (pmx > 0.0f) atomicMax(&a.unit_max[1], __float_as_uint(pmx));
        if (vmx > 0.0f) atomicMax(&a.unit_max[2], __float_as_uint(vmx));
    }
}

}  // namespace sygdev

namespace syglaunch {

int dct_matrix(const double* S, const double* D, long long n_units, int N, long long T, int C, double* out, int sm_count, cudaStream_t st,
               std::string& err) {
    const long long total = n_units * C * T;
    if (total <= 0) return 0;
    const int grid = (int)std::min<long long>((total + sygdev::kThreads - 1) / sygdev::kThreads, (long long)sm_count * 16);
    SYG_LAUNCH(sygdev::dct_matrix_kernel, grid, sygdev::kThreads, 0, st, S, D, n_units, N, T, C, out);
    LCK(cudaGetLastError());
    return 0;
}

int contrast_spectrum(const float* S, int B, long long T, int nb, const int* lo, const int* cnt, const int* nq, float* cws, unsigned* unit_max,
                      int sm_count, cudaStream_t st, std::string& err) {
    if (T <= 0) return 0;
    sygdev::ContrastSpecArgs a;
    a.S = S; a.B = B; a.T = T; a.nb = nb; a.cws = cws; a.unit_max = unit_max;
    for (int i = 0; i < syg::kMaxBands; ++i) { a.band_lo[i] = i < nb ? lo[i] : 0; a.band_cnt[i] = i < nb ? cnt[i] : 0; a.band_n[i] = i < nb ? nq[i] : 1; }
    a.pw = ((B + 4 * (B >> 5) + 48 + 3) / 4) * 4;
    const size_t smem = (size_t)(sygdev::kThreads / 32) * a.pw * sizeof(float);
    if (smem > 227 * 1024) { err = "spectral_contrast(S=...): more than 4097 frequency rows are not supported"; return -5; }
    static KernelCache kc;
    int bps = 0;
    if (int rc = prepare_kernel(sygdev::contrast_spectrum_kernel, sygdev::kThreads, smem, kc, &bps, err)) return rc;
    const long long want = (T + sygdev::kThreads / 32 - 1) / (sygdev::kThreads / 32);
    const int grid = (int)std::min<long long>(want, (long long)sm_count * bps);
    SYG_LAUNCH(sygdev::contrast_spectrum_kernel, grid, sygdev::kThreads, smem, st, a);
    LCK(cudaGetLastError());
    return 0;
}

}  // namespace syglaunch
