// Welch / periodogram kernel instantiations
#include "syg_launch_common.h"
#include "syg_kernels.cuh"

namespace syglaunch {

template <class TL>
static int welch_t(const syg::WelchArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using SM = sygdev::WelchSmem<TL>;
    static int blocks_per_sm = 0;
    auto kfn = sygdev::welch_kernel<TL>;
    if (blocks_per_sm == 0) {
        LCK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes));
        int nb = 0;
        LCK(SYG_OCCUPANCY(nb, kfn, sygdev::kThreads, SM::bytes));
        if (nb < 1) { err = "welch kernel does not fit on an SM"; return -3; }
        blocks_per_sm = nb;
    }
    if (a.g.n_units <= 0) return 0;
    const int grid = (int)std::min<long long>(a.g.n_units, (long long)sm_count * blocks_per_sm);
    SYG_LAUNCH(kfn, grid, sygdev::kThreads, SM::bytes, st, a);
    LCK(cudaGetLastError());
    return 0;
}

int welch(int nfft, const syg::WelchArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    using namespace sygdev;
    switch (ilog2i(nfft / 2)) {
        case 4: return welch_t<FftTile<4, 4>>(a, sm_count, st, err);
        case 5: return welch_t<FftTile<5, 8>>(a, sm_count, st, err);
        case 6: return welch_t<FftTile<6, 8>>(a, sm_count, st, err);
        case 7: return welch_t<FftTile<7, 16>>(a, sm_count, st, err);
        case 8: return welch_t<FftTile<8, 16>>(a, sm_count, st, err);
        case 9: return welch_t<FftTile<9, 16>>(a, sm_count, st, err);
        case 10: return welch_t<FftTile<10, 16>>(a, sm_count, st, err);
        case 11: return welch_t<FftTile<11, 16>>(a, sm_count, st, err);
        case 12: return welch_t<FftTile<12, 16>>(a, sm_count, st, err);
    }
    err = "nfft=" + std::to_string(nfft) + ": only powers of two in [32, 8192] are supported";
    return -5;
}

}  // namespace syglaunch
