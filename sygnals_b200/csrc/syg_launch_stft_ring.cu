// STFT magnitude / power through the TMA-staged ring kernel (syg_stft_ring.cuh): compute_stft (dsp.py:167-229), n_fft 256..2048
#include <cstdlib>

#include "syg_launch_common.h"
#include "syg_stft_ring.cuh"

namespace syglaunch {

template <class TL, int NW, bool DB, int S>
static int stft_ring_t(const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using RG = sygdev::RingGeom<TL, NW, DB>;
    auto kfn = sygdev::stft_ring_kernel<TL, NW, DB, S>;
    const size_t smem = RG::bytes(a.hop, S);
    if (smem > 227 * 1024) return 1;                                 // hop too large for the ring: the caller takes the other kernel
    static KernelCache kc;
    int bps = 0;
    if (int rc = prepare_kernel(kfn, RG::NT, smem, kc, &bps, err)) return rc;
    const long long n_rounds = a.g.n_units * (long long)((a.T + RG::TT - 1) / RG::TT);
    if (n_rounds <= 0) return 0;
    if (n_rounds >= (1LL << 31) || a.g.n_units >= (1LL << 31)) return 1;            // the kernel numbers rounds in 32 bits
    const int grid = (int)std::min<long long>(n_rounds, (long long)sm_count * bps);
    SYG_LAUNCH(kfn, grid, RG::NT, smem, st, a);
    LCK(cudaGetLastError());
    return 0;
}

// returns 0 (launched), 1 (not eligible: use the register-staged kernels) or a negative error
int stft_ring(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using namespace sygdev;
    static int mode = -2;                                             // SYGB200_RING=0: off (A/B against the register-staged kernels)
    if (mode == -2) { const char* e = std::getenv("SYGB200_RING"); mode = e ? std::atoi(e) : 1; }
    if (mode == 0) return 1;
    // eligibility: real output, zero padding, float2-aligned frames, 16-byte aligned bulk copies for every round of every unit
    if (a.out_kind == 0 || a.pad_mode != 0 || (a.hop & 1) || a.T <= 0) return 1;
    if (a.g.unit_starts || a.g.unit_valid) return 1;
    if ((reinterpret_cast<uintptr_t>(a.y) & 15u) || (a.g.unit_stride & 3) || (a.g.unit_len & 3) || (a.g.total_len & 3) || (a.g.unit0 != 0 && (a.g.unit_stride & 3)))
        return 1;
    switch (n_fft) {
        case 256: return stft_ring_t<FftTile<7, 16>, 16, true, 2>(a, sm_count, st, err);
        case 512: return stft_ring_t<FftTile<8, 16>, 16, true, 2>(a, sm_count, st, err);
        case 1024: return stft_ring_t<FftTile<9, 32>, 8, true, 2>(a, sm_count, st, err);
        // measured on B200 (cfg2, round 2): 20 warps for n_fft 512 (no L1 left) 0.485 vs 0.382 ms; single tile + two barriers with
        // 12 / 10 warps for 1024 / 2048: 0.465 / 0.621 vs 0.471 / 0.489 ms; three stages instead of two: no change
        case 2048: return stft_ring_t<FftTile<10, 32>, 8, true, 2>(a, sm_count, st, err);
    }
    return 1;
}

}  // namespace syglaunch
