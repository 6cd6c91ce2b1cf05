// warp-synchronous feature kernel instantiations (n_fft <= 2048), fused, EXTRA=true
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_extra(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_dispatch<true, 0>(n_fft, a, sm_count, st, err);
}
int frame_warp_base(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err);
int frame_warp_s1(int n_fft, bool extra, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err);
int frame_warp_s2(int n_fft, bool extra, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err);
// stage: 0 fused, 1 / 2 the two-stage launch
int frame_warp(int n_fft, bool extra, int stage, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    if (stage == 1) return frame_warp_s1(n_fft, extra, a, sm_count, st, err);
    if (stage == 2) return frame_warp_s2(n_fft, extra, a, sm_count, st, err);
    return extra ? frame_warp_extra(n_fft, a, sm_count, st, err) : frame_warp_base(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
