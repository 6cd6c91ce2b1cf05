// warp-synchronous feature kernel instantiations (n_fft <= 2048), EXTRA=true
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_extra(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_dispatch<true, 0>(n_fft, a, sm_count, st, err);
}
int frame_warp_base(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err);
int frame_warp(int n_fft, bool extra, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return extra ? frame_warp_extra(n_fft, a, sm_count, st, err) : frame_warp_base(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
