// Shared device/host helpers for libdmt_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dmt_b200.h"

namespace dmt {

void set_error(const char* msg);

#define DMT_CUDA(expr)                        \
    do {                                      \
        cudaError_t _e = (expr);              \
        if (_e != cudaSuccess) {              \
            dmt::set_error(cudaGetErrorString(_e)); \
            return (int)_e;                   \
        }                                     \
    } while (0)

#define DMT_REQUIRE(cond, msg)     \
    do {                           \
        if (!(cond)) {             \
            dmt::set_error(msg);   \
            return DMT_E_ARG;      \
        }                          \
    } while (0)

void count_launch(long long n);
long long launch_count();

#define DMT_LAUNCH_CHECK()              \
    do {                                \
        dmt::count_launch(1);           \
        DMT_CUDA(cudaGetLastError());   \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// A batch of rows inside an epoch plan. When row_off == nullptr the range is [lo, hi) by value (stateless
// C-ABI calls); otherwise it is read from device memory so that a captured CUDA graph can be replayed for
// a different permutation: rows [row_off[b], row_off[b+1]) and `active[b]` gates the whole step.
struct BatchRef {
    const int32_t* row_off;
    const int32_t* active;
    int b;
    int lo, hi;
};

__device__ __forceinline__ bool batch_range(const BatchRef& r, int& lo, int& hi) {
    if (r.row_off == nullptr) {
        lo = r.lo;
        hi = r.hi;
        return true;
    }
    if (r.active != nullptr && r.active[r.b] == 0) return false;
    lo = r.row_off[r.b];
    hi = r.row_off[r.b + 1];
    return true;
}

inline BatchRef batch_by_value(int lo, int hi) { return BatchRef{nullptr, nullptr, 0, lo, hi}; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a block of up to 1024 threads; result valid in every thread. `sh` needs 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (lane < nw) ? sh[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Per-element loss and its derivative w.r.t. the prediction (reference src/models/utils.py:7-14).
__device__ __forceinline__ float loss_value(int kind, float o, float y) {
    if (kind == DMT_LOSS_MSE) {
        float d = o - y;
        return d * d;
    }
    // BCE with logits, stable form: max(o,0) - o*y + log1p(exp(-|o|))
    return fmaxf(o, 0.f) - o * y + log1pf(expf(-fabsf(o)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float loss_grad(int kind, float o, float y) {
    if (kind == DMT_LOSS_MSE) return 2.f * (o - y);
    return sigmoidf_(o) - y;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL): a kernel launched with `pdl` may have its blocks scheduled while the previous
// kernel of the stream is still draining; DMT_PDL_ENTRY() at the top of the kernel first lets the NEXT kernel do the
// same and then waits until the previous kernel has completed and flushed its memory, so every load / store of the
// body is ordered exactly as without PDL. What overlaps is launch latency and block scheduling (~2 us per edge of the
// six-kernel step chain). Without the attribute both calls are no-ops.
#define DMT_PDL_ENTRY()                            \
    do {                                           \
        cudaTriggerProgrammaticLaunchCompletion(); \
        cudaGridDependencySynchronize();           \
    } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace dmt
