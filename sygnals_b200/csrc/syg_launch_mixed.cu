// Mixed-radix kernel instantiations: features / STFT / Welch for transform lengths that are not powers of two (syg_mixed.cuh).
#include "syg_launch_common.h"
#include "syg_mixed.cuh"

namespace syglaunch {

template <int MODE>
static int frame_mixed_t(const syg::FrameArgs& a, const syg::MixedPlan& mp, int sm_count, cudaStream_t st, std::string& err) {
    auto kfn = sygdev::frame_mixed_kernel<MODE>;
    const size_t smem = sygdev::mixed_layout(mp.L, mp.B, MODE == sygdev::MODE_FEATURES).bytes;
    static KernelCache kc;
    int blocks_per_sm = 0;
    if (int rc = prepare_kernel(kfn, sygdev::kThreads, smem, kc, &blocks_per_sm, err)) return rc;
    if (a.n_frames <= 0) return 0;
    const int grid = (int)std::min<long long>(a.n_frames, (long long)sm_count * blocks_per_sm);
    SYG_LAUNCH(kfn, grid, sygdev::kThreads, smem, st, a, mp);
    LCK(cudaGetLastError());
    return 0;
}

int frame_mixed(int mode, const syg::FrameArgs& a, const syg::MixedPlan& mp, int sm_count, cudaStream_t st, std::string& err) {
    return mode == sygdev::MODE_STFT ? frame_mixed_t<sygdev::MODE_STFT>(a, mp, sm_count, st, err)
                                     : frame_mixed_t<sygdev::MODE_FEATURES>(a, mp, sm_count, st, err);
}

int welch_mixed(const syg::WelchArgs& a, const syg::MixedPlan& mp, int sm_count, cudaStream_t st, std::string& err) {
    auto kfn = sygdev::welch_mixed_kernel;
    const size_t smem = sygdev::mixed_layout(mp.L, mp.B, false).bytes;
    static KernelCache kc;
    int blocks_per_sm = 0;
    if (int rc = prepare_kernel(kfn, sygdev::kThreads, smem, kc, &blocks_per_sm, err)) return rc;
    if (a.g.n_units <= 0) return 0;
    const int grid = (int)std::min<long long>(a.g.n_units, (long long)sm_count * blocks_per_sm);
    SYG_LAUNCH(kfn, grid, sygdev::kThreads, smem, st, a, mp);
    LCK(cudaGetLastError());
    return 0;
}

}  // namespace syglaunch
