// dispatch of the warp-synchronous feature kernel over n_fft (included by the two warp translation units)
#pragma once

#include <cstdlib>

#include "syg_launch_common.h"
#include "syg_frame_warp.cuh"

namespace syglaunch {

template <class TL, bool EXTRA, int NT = sygdev::kThreads, int MINB = 2, bool SYNCP = false>
static int frame_warp_t(const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using WT = sygdev::WarpTile<TL, NT>;
    static int blocks_per_sm = 0;
    auto kfn = sygdev::frame_warp_kernel<TL, EXTRA, NT, MINB, SYNCP>;
    if (blocks_per_sm == 0) {
        LCK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT::bytes));
        int nb = 0;
        LCK(SYG_OCCUPANCY(nb, kfn, NT, WT::bytes));
        if (nb < 1) { err = "frame_warp kernel does not fit on an SM"; return -3; }
        blocks_per_sm = nb;
    }
    const long long per_cta = (long long)WT::FW * WT::kWarps;
    const long long n_rounds = (a.n_frames + per_cta - 1) / per_cta;
    if (n_rounds <= 0) return 0;
    const int grid = (int)std::min<long long>(n_rounds, (long long)sm_count * blocks_per_sm);
    SYG_LAUNCH(kfn, grid, NT, WT::bytes, st, a);
    LCK(cudaGetLastError());
    return 0;
}

template <bool EXTRA>
static int frame_warp_dispatch(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using namespace sygdev;
    switch (ilog2i(n_fft / 2)) {
        case 4: return frame_warp_t<FftTile<4, 4>, EXTRA>(a, sm_count, st, err);
        case 5: return frame_warp_t<FftTile<5, 8>, EXTRA>(a, sm_count, st, err);
        case 6: return frame_warp_t<FftTile<6, 8>, EXTRA>(a, sm_count, st, err);
        case 7: return frame_warp_t<FftTile<7, 16>, EXTRA>(a, sm_count, st, err);
        case 8: return frame_warp_t<FftTile<8, 16>, EXTRA>(a, sm_count, st, err);
        case 9: return frame_warp_t<FftTile<9, 32>, EXTRA>(a, sm_count, st, err);
        case 10: {
            if (!EXTRA) {       // tuning variants (SYGB200_VARIANT), see DESIGN.md
                static int variant = -1;
                if (variant < 0) { const char* e = std::getenv("SYGB200_VARIANT"); variant = e ? std::atoi(e) : 0; }
                if (variant == 1) return frame_warp_t<FftTile<10, 32>, false, 256, 2, true>(a, sm_count, st, err);
                if (variant == 2) return frame_warp_t<FftTile<10, 32>, false, 512, 1, true>(a, sm_count, st, err);
                if (variant == 3) return frame_warp_t<FftTile<10, 32>, false, 512, 1, false>(a, sm_count, st, err);
                if (variant == 4) return frame_warp_t<FftTile<10, 32>, false, 320, 2, false>(a, sm_count, st, err);
                if (variant == 5) return frame_warp_t<FftTile<10, 32>, false, 384, 2, false>(a, sm_count, st, err);
            }
            return frame_warp_t<FftTile<10, 32>, EXTRA>(a, sm_count, st, err);
        }
    }
    err = "n_fft=" + std::to_string(n_fft) + " has no warp tile";
    return -5;
}

}  // namespace syglaunch
