// n_fft 2048 feature kernel specialised for the 22.05 kHz plan (Spec22k: the reference's default sample rate)
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_spec22k(const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_t<sygdev::FftTile<10, 32>, false, SYG_NT2048, 1, 0, sygdev::Spec22k>(a, sm_count, st, err);
}
}  // namespace syglaunch
