// warp-synchronous feature kernel instantiations (n_fft <= 2048), EXTRA=false
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_base(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_dispatch<false, 0>(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
