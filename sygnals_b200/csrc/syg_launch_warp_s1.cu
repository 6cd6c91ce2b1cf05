// two-stage launch, stage 1 (framing + FFT + |X|^2 -> spectra workspace)
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_s1(int n_fft, bool extra, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return extra ? frame_warp_dispatch<true, 1>(n_fft, a, sm_count, st, err) : frame_warp_dispatch<false, 1>(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
