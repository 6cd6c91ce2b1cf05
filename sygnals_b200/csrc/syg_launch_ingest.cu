// PCM16 ingest: int16 samples -> float32 in [-1, 1) exactly as libsndfile / soundfile.read(dtype='float32') normalises them
// (x / 32768), which is what librosa.load hands the reference (sygnals/core/audio/io.py:84-95).
#include "syg_launch_common.h"
#include "syg_device.cuh"

namespace sygdev {
__global__ void __launch_bounds__(kThreads) pcm16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n8 = n >> 3;                                    // 8 samples (16 bytes in, 32 bytes out) per thread step
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    long long done = 0;
    if (aligned) {
        const int4* in4 = reinterpret_cast<const int4*>(in);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n8; i += stride) {
            const int4 v = __ldg(in4 + i);
            const int w[4] = {v.x, v.y, v.z, v.w};
            float f[8];
            SYG_UNROLL
            for (int k = 0; k < 4; ++k) {
                f[2 * k] = (float)(short)(w[k] & 0xffff) * (1.0f / 32768.0f);
                f[2 * k + 1] = (float)(short)((unsigned)w[k] >> 16) * (1.0f / 32768.0f);
            }
            out4[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
            out4[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
        }
        done = n8 << 3;
    }
    for (long long i = done + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
        out[i] = (float)in[i] * (1.0f / 32768.0f);
}
}  // namespace sygdev

namespace syglaunch {
int pcm16_to_f32(const short* in, float* out, long long n, int sm_count, cudaStream_t st, std::string& err) {
    if (n <= 0) return 0;
    const long long want = (n / 8 + sygdev::kThreads - 1) / sygdev::kThreads + 1;
    const int grid = (int)std::min<long long>(want, (long long)sm_count * 8);
    SYG_LAUNCH(sygdev::pcm16_to_f32_kernel, grid, sygdev::kThreads, 0, st, in, out, n);
    LCK(cudaGetLastError());
    return 0;
}
}  // namespace syglaunch
