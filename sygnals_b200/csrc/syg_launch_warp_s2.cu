// two-stage launch, stage 2 (spectra workspace -> statistics, mel, contrast), EXTRA=false
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_s2x(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err);
int frame_warp_s2(int n_fft, bool extra, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return extra ? frame_warp_s2x(n_fft, a, sm_count, st, err) : frame_warp_dispatch<false, 2>(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
