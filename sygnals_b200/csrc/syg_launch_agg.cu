// Segment-level aggregation of frame features on the device: the consumer of the gathered feature matrix in
// `sygnals save dataset` is format_feature_vectors_per_segment (sygnals/core/ml_utils/formatters.py:51-163), a NaN-aware
// mean / std / median / min / max over each segment's frames (formatters.py:28-47).  Doing it here shrinks the final gather
// from [segments, rows, T] to [segments, rows].
#include "syg_launch_common.h"
#include "syg_device.cuh"

namespace sygdev {

struct AggArgs {
    const float* feats;
    long long n_seg;
    int n_rows;
    long long row_stride;          // elements between consecutive rows of one segment
    const long long* seg_off;      // optional [n_seg]: first element of row 0 of the segment (NULL: s * n_rows * row_stride)
    const int* seg_len;            // optional [n_seg]: frames of the segment (NULL: fixed_len); <= 0 -> NaN row (skipped segment)
    int fixed_len;
    int agg[64];                   // per row: 0 mean, 1 std, 2 median, 3 min, 4 max
    double* out;                   // [n_seg][n_rows]
};

SYG_DEVICE SYG_INLINE unsigned ordered_key(float x) {               // order preserving float -> unsigned
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SYG_DEVICE SYG_INLINE float key_value(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
SYG_DEVICE SYG_INLINE double warp_sum_d(double v) {
    SYG_UNROLL
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// key of the k-th smallest (0-based) non-NaN value of x[0..len): bitwise search from the most significant bit
SYG_DEVICE SYG_INLINE unsigned warp_kth_key(const float* __restrict__ x, int len, int k, int lane) {
    unsigned prefix = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned trial = prefix | (1u << bit);
        int cnt = 0;
        for (int i = lane; i < len; i += 32) {
            const float v = x[i];
            cnt += (v == v && ordered_key(v) < trial) ? 1 : 0;
        }
        cnt = __reduce_add_sync(kFull, cnt);
        if (cnt <= k) prefix = trial;
    }
    return prefix;
}

__global__ void __launch_bounds__(kThreads) aggregate_kernel(const AggArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kThreads / 32);
    const long long n_items = a.n_seg * a.n_rows;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long item = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); item < n_items; item += warps) {
        const long long s = item / a.n_rows;
        const int r = (int)(item - s * a.n_rows);
        const int len = a.seg_len ? a.seg_len[s] : a.fixed_len;
        double res = qnan;
        if (len > 0) {
            const long long base = (a.seg_off ? a.seg_off[s] : s * a.n_rows * a.row_stride) + (long long)r * a.row_stride;
            const float* x = a.feats + base;
            int n = 0;
            double sum = 0.0;
            float mn = __uint_as_float(0x7f800000u), mx = __uint_as_float(0xff800000u);
            // eight loads in flight per lane before the first use (a warp walks ~90 items one after the other: the dependent
            // load -> FP64 add chain of the plain loop left the kernel latency bound); same elements in the same order per lane
            const float fnan = __uint_as_float(0x7fc00000u);
            for (int i0 = lane; i0 < len; i0 += 256) {
                float v[8];
                SYG_UNROLL
                for (int k = 0; k < 8; ++k) v[k] = (i0 + 32 * k < len) ? x[i0 + 32 * k] : fnan;
                SYG_UNROLL
                for (int k = 0; k < 8; ++k)
                    if (v[k] == v[k]) { ++n; sum += (double)v[k]; mn = fminf(mn, v[k]); mx = fmaxf(mx, v[k]); }
            }
            n = __reduce_add_sync(kFull, n);
            if (n > 0) {                                            // formatters.py:31-36: all-NaN -> NaN
                const int kind = a.agg[r];
                if (kind == 0 || kind == 1) {
                    const double mean = warp_sum_d(sum) / (double)n;
                    res = mean;
                    if (kind == 1) {                                // np.std: population (ddof = 0), two passes
                        double ss = 0.0;
                        for (int i0 = lane; i0 < len; i0 += 256) {
                            float v[8];
                            SYG_UNROLL
                            for (int k = 0; k < 8; ++k) v[k] = (i0 + 32 * k < len) ? x[i0 + 32 * k] : fnan;
                            SYG_UNROLL
                            for (int k = 0; k < 8; ++k)
                                if (v[k] == v[k]) { const double d = (double)v[k] - mean; ss += d * d; }
                        }
                        res = sqrt(warp_sum_d(ss) / (double)n);
                    }
                } else if (kind == 3) {
                    SYG_UNROLL
                    for (int o = 16; o >= 1; o >>= 1) mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
                    res = (double)mn;
                } else if (kind == 4) {
                    SYG_UNROLL
                    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
                    res = (double)mx;
                } else {                                            // np.median: middle value, or the mean of the two middle ones
                    const unsigned khi = warp_kth_key(x, len, n / 2, lane);
                    res = (double)key_value(khi);
                    if ((n & 1) == 0) {
                        const unsigned klo = warp_kth_key(x, len, n / 2 - 1, lane);
                        res = 0.5 * ((double)key_value(klo) + res);
                    }
                }
            }
        }
        if (lane == 0) a.out[item] = res;
    }
}

}  // namespace sygdev

namespace syglaunch {
int aggregate(const float* feats, long long n_seg, int n_rows, long long row_stride, const long long* seg_off, const int* seg_len,
              int fixed_len, const int* agg_host, double* out, int sm_count, cudaStream_t st, std::string& err) {
    if (n_seg <= 0 || n_rows <= 0) return 0;
    if (n_rows > 64) { err = "aggregation supports at most 64 rows per call"; return -5; }
    sygdev::AggArgs a;
    a.feats = feats; a.n_seg = n_seg; a.n_rows = n_rows; a.row_stride = row_stride; a.seg_off = seg_off; a.seg_len = seg_len;
    a.fixed_len = fixed_len; a.out = out;
    for (int i = 0; i < 64; ++i) a.agg[i] = i < n_rows ? agg_host[i] : 0;
    const long long items = n_seg * n_rows;
    const int grid = (int)std::min<long long>((items + sygdev::kThreads / 32 - 1) / (sygdev::kThreads / 32), (long long)sm_count * 8);
    SYG_LAUNCH(sygdev::aggregate_kernel, grid, sygdev::kThreads, 0, st, a);
    LCK(cudaGetLastError());
    return 0;
}
}  // namespace syglaunch
