// STFT output of the warp-synchronous kernel (n_fft <= 2048): stage 3 = framing + FFT + real split -> transposed CTA tile ->
// contiguous row stores  (dsp.py:167-229 compute_stft; complex64 / magnitude / power)
#include "syg_launch_warp.h"

namespace syglaunch {
int frame_warp_stft(int n_fft, const syg::FrameArgs& a, int sm_count, cudaStream_t st, std::string& err) {
    return frame_warp_dispatch<false, 3>(n_fft, a, sm_count, st, err);
}
}  // namespace syglaunch
