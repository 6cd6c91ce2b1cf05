// STFT magnitude / power for n_fft 4096 / 8192 through stft_big_kernel (syg_stft_big.cuh): compute_stft (dsp.py:167-229)
#include "syg_launch_common.h"
#include "syg_stft_big.cuh"

namespace syglaunch {

template <int R, int NW>
static int stft_big_t(const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using BG = sygdev::BigGeom<R, NW>;
    auto kfn = sygdev::stft_big_kernel<R, NW>;
    static KernelCache kc;
    int bps = 0;
    if (int rc = prepare_kernel(kfn, BG::NT, BG::bytes, kc, &bps, err)) return rc;
    const long long n_rounds = (a.n_frames + BG::TT - 1) / BG::TT;
    if (n_rounds <= 0) return 0;
    const int grid = (int)std::min<long long>(n_rounds, (long long)sm_count * bps);
    SYG_LAUNCH(kfn, grid, BG::NT, BG::bytes, st, a);
    LCK(cudaGetLastError());
    return 0;
}

int stft_big(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    if (n_fft == 4096) return stft_big_t<2, 16>(a, sm_count, st, err);      // 16 warps, 8 frames per round: 8 x 2 x 8.3 KB regions + 72 KB tile
    if (n_fft == 8192) return stft_big_t<4, 16>(a, sm_count, st, err);      // 16 warps, 4 frames per round: 4 x 4 x 8.3 KB regions + 80 KB tile
    err = "stft_big: n_fft must be 4096 or 8192";
    return -5;
}

}  // namespace syglaunch
