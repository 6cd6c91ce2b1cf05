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

// General ingest (load_audio(mono=True), sygnals/core/audio/io.py:38-102 -> librosa.load -> soundfile.read(dtype=float32) + to_mono):
// interleaved PCM frames of `channels` samples -> one float32 per frame.  Normalisation as libsndfile does it (u8: (x - 128) / 128,
// s16: x / 2^15, s24: x / 2^23, s32: x / 2^31, f32: unchanged); the channel mean as numpy computes np.mean(y, axis=0) on a float32
// array: float32 sum over the channels in numpy's order (sequential below 8 channels, its 8-way pairwise scheme from 8 on), then one
// float32 division by the count.  FMT: 0 u8, 1 s16, 2 s24, 3 s32, 4 f32.
template <int FMT>
SYG_DEVICE SYG_INLINE float pcm_sample(const unsigned char* __restrict__ p) {
    if (FMT == 0) return ((float)(int)__ldg(p) - 128.0f) * (1.0f / 128.0f);
    if (FMT == 1) return (float)__ldg(reinterpret_cast<const short*>(p)) * (1.0f / 32768.0f);
    if (FMT == 2) {
        const int v = (int)__ldg(p) | ((int)__ldg(p + 1) << 8) | ((int)(signed char)__ldg(p + 2) << 16);   // little endian, sign from the top byte
        return (float)v * (1.0f / 8388608.0f);
    }
    if (FMT == 3) return (float)__ldg(reinterpret_cast<const int*>(p)) * (1.0f / 2147483648.0f);
    return __ldg(reinterpret_cast<const float*>(p));
}

template <int FMT>
__global__ void __launch_bounds__(kThreads) pcm_ingest_kernel(const unsigned char* __restrict__ raw, int channels, long long n_frames,
                                                              float* __restrict__ out) {
    constexpr int BPS = FMT == 0 ? 1 : (FMT == 1 ? 2 : (FMT == 2 ? 3 : 4));
    const long long stride = (long long)gridDim.x * kThreads;
    const float cnt = (float)channels;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_frames; i += stride) {
        const unsigned char* p = raw + i * (long long)(channels * BPS);
        float acc;
        if (channels < 8) {
            acc = pcm_sample<FMT>(p);
            for (int c = 1; c < channels; ++c) acc += pcm_sample<FMT>(p + c * BPS);
        } else {
            // numpy's pairwise_sum along the contiguous axis (the channels of one frame after librosa's transpose): eight running
            // sums over blocks of eight, combined as a tree, then the tail added one by one (channels <= 64 < its 128 block size)
            float r[8];
            SYG_UNROLL
            for (int k = 0; k < 8; ++k) r[k] = pcm_sample<FMT>(p + k * BPS);
            int c = 8;
            for (; c + 8 <= channels; c += 8) {
                SYG_UNROLL
                for (int k = 0; k < 8; ++k) r[k] += pcm_sample<FMT>(p + (c + k) * BPS);
            }
            acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; c < channels; ++c) acc += pcm_sample<FMT>(p + c * BPS);
        }
        out[i] = channels > 1 ? acc / cnt : acc;
    }
}
}  // namespace sygdev

namespace syglaunch {
int pcm16_to_f32(const short* in, float* out, long long n, int sm_count, cudaStream_t st, std::string& err);
int pcm_to_f32(const void* raw, int fmt, int channels, long long n_frames, float* out, int sm_count, cudaStream_t st, std::string& err) {
    if (n_frames <= 0) return 0;
    if (channels < 1 || channels > 64) { err = "channels must be in [1, 64]"; return -1; }
    if (fmt == 1 && channels == 1) return pcm16_to_f32(reinterpret_cast<const short*>(raw), out, n_frames, sm_count, st, err);   // vectorised
    const unsigned char* r = reinterpret_cast<const unsigned char*>(raw);
    const int grid = (int)std::min<long long>((n_frames + sygdev::kThreads - 1) / sygdev::kThreads, (long long)sm_count * 16);
    switch (fmt) {
        case 0: SYG_LAUNCH(sygdev::pcm_ingest_kernel<0>, grid, sygdev::kThreads, 0, st, r, channels, n_frames, out); break;
        case 1: SYG_LAUNCH(sygdev::pcm_ingest_kernel<1>, grid, sygdev::kThreads, 0, st, r, channels, n_frames, out); break;
        case 2: SYG_LAUNCH(sygdev::pcm_ingest_kernel<2>, grid, sygdev::kThreads, 0, st, r, channels, n_frames, out); break;
        case 3: SYG_LAUNCH(sygdev::pcm_ingest_kernel<3>, grid, sygdev::kThreads, 0, st, r, channels, n_frames, out); break;
        case 4: SYG_LAUNCH(sygdev::pcm_ingest_kernel<4>, grid, sygdev::kThreads, 0, st, r, channels, n_frames, out); break;
        default: err = "unknown PCM sample format " + std::to_string(fmt); return -1;
    }
    LCK(cudaGetLastError());
    return 0;
}

int pcm16_to_f32(const short* in, float* out, long long n, int sm_count, cudaStream_t st, std::string& err) {
    if (n <= 0) return 0;
    const long long want = (n / 8 + sygdev::kThreads - 1) / sygdev::kThreads + 1;
    const int grid = (int)std::min<long long>(want, (long long)sm_count * 8);
    SYG_LAUNCH(sygdev::pcm16_to_f32_kernel, grid, sygdev::kThreads, 0, st, in, out, n);
    LCK(cudaGetLastError());
    return 0;
}
}  // namespace syglaunch
