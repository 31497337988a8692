// Privacy transforms of the broadcast pseudo-residuals on the device (reference src/privacy.py:6-58, applied at
// src/assist.py:59-60): both clip the vector to its [2.5 %, 97.5 %] quantile range [a, b] first.
//   dp : y' = clip(y, a, b) + Laplace(scale = (b - a) / alpha)
//   ip : y' = (1/T) * sum over T uniform thresholds t in [a, b] of (2t - b if y < t else 2t - a)
// The quantiles follow numpy's default (linear interpolation between order statistics), computed from a radix sort of
// the keys. The noise comes from a counter-based generator keyed on (seed, element, draw): reproducible and
// independent of the launch geometry — the production-mode replacement of numpy's global generator, which only the
// host path (privacy.py, rng="reference") can replay draw for draw.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace dmt {

__global__ void quantile_pair_kernel(const float* __restrict__ sorted, int64_t n, float q_lo, float q_hi,
                                     float* __restrict__ out) {
    if (threadIdx.x >= 2 || blockIdx.x != 0) return;
    const double q = threadIdx.x == 0 ? (double)q_lo : (double)q_hi;
    // numpy 'linear': virtual index q * (n - 1), interpolate between its neighbours
    const double pos = q * (double)(n - 1);
    int64_t i = (int64_t)floor(pos);
    if (i < 0) i = 0;
    if (i > n - 1) i = n - 1;
    const int64_t j = i + 1 < n ? i + 1 : i;
    const double g = pos - (double)i;
    const double a = (double)sorted[i], b = (double)sorted[j];
    // numpy's _lerp: a + (b - a) * g, mirrored for g >= 0.5 to stay monotone
    const double v = g >= 0.5 ? b - (b - a) * (1.0 - g) : a + (b - a) * g;
    out[threadIdx.x] = (float)v;
}

__device__ __forceinline__ float unit_open(uint32_t h) {  // (0, 1)
    return ((float)(h >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256) privacy_kernel(const float* __restrict__ y, int64_t n,
                                                      const float* __restrict__ ab, int mode, float param,
                                                      uint64_t seed, float* __restrict__ out) {
    const float a = ab[0], b = ab[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = y[i];
        if (mode == 0) {  // dp
            const float scale = fmaxf(0.f, (b - a) / param);
            const float u = unit_open(mix32(seed ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL))) - 0.5f;  // (-0.5, 0.5)
            // inverse CDF of Laplace(0, scale)
            const float noise = -scale * copysignf(1.f, u) * log1pf(-2.f * fabsf(u));
            out[i] = fminf(fmaxf(v, a), b) + noise;
        } else {  // ip: `param` thresholds per element
            const int T = (int)param;
            float acc = 0.f;
            for (int t = 0; t < T; ++t) {
                const float u = unit_open(mix32(seed ^ ((uint64_t)i * 0x9E3779B97F4A7C15ULL) ^
                                                ((uint64_t)(t + 1) * 0xD6E8FEB86659FD93ULL)));
                const float thr = a + (b - a) * u;
                acc += (v < thr ? 2.f * thr - b : 2.f * thr - a) / (float)T;
            }
            out[i] = acc;
        }
    }
}

}  // namespace dmt

using namespace dmt;

static int64_t sorted_area_bytes(int64_t n) { return (n * 4 + 64 + 255) / 256 * 256; }  // keeps CUB's area 256 B aligned

extern "C" {

int64_t dmt_privacy_temp_bytes(int64_t n) {
    if (n <= 0) return 16;
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const float*)nullptr, (float*)nullptr, n);
    return (int64_t)bytes + sorted_area_bytes(n);  // sort temp + the sorted copy + the quantile pair
}

int dmt_privacy(const float* y, int64_t n, int mode, float param, uint64_t seed, float* out, float* quantiles,
                void* temp, int64_t temp_bytes, void* stream) {
    DMT_REQUIRE(y && out && n >= 0 && (mode == 0 || mode == 1) && param > 0.f, "dmt_privacy: bad argument");
    DMT_REQUIRE(n < (1LL << 31), "dmt_privacy: n must fit int32 (radix sort item count)");
    if (n == 0) return 0;
    DMT_REQUIRE(temp && temp_bytes >= dmt_privacy_temp_bytes(n), "dmt_privacy: temp too small");
    cudaStream_t st = as_stream(stream);
    float* sorted = reinterpret_cast<float*>(temp);
    float* ab = sorted + n;
    void* cub_temp = reinterpret_cast<char*>(temp) + sorted_area_bytes(n);
    size_t cub_bytes = (size_t)(temp_bytes - sorted_area_bytes(n));
    DMT_CUDA(cub::DeviceRadixSort::SortKeys(cub_temp, cub_bytes, y, sorted, (int)n, 0, 32, st));
    quantile_pair_kernel<<<1, 32, 0, st>>>(sorted, n, 0.025f, 0.975f, ab);
    DMT_LAUNCH_CHECK();
    int blocks = (int)((n + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    privacy_kernel<<<blocks, 256, 0, st>>>(y, n, ab, mode, param, seed, out);
    DMT_LAUNCH_CHECK();
    if (quantiles) DMT_CUDA(cudaMemcpyAsync(quantiles, ab, 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // extern "C"
