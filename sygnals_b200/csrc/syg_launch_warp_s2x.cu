// two-stage launch, stage 2, EXTRA=true
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_s2x(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_dispatch<true, 2>(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
