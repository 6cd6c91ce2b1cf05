// n_fft 2048 feature kernel specialised for the 44.1 kHz plan AND the feature set / output rows of BASELINE cfg4 (Spec44kL)
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_spec44kl(const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_t<sygdev::FftTile<10, 32>, false, SYG_NT2048, 1, 0, sygdev::Spec44kL>(a, sm_count, st, err);
}
}  // namespace syglaunch
