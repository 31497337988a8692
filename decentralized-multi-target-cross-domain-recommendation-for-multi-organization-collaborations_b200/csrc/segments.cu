// Sort-by-index and warp-segmented reduction: the dense-gradient form of autograd's embedding / gather backward
// without contended atomics (reference call sites: nn.Embedding at src/models/mf.py:37,44; weight gathers at
// src/models/ae.py:102,135-136; torch.sort/unique_consecutive at ae.py:103-104,137-139).
//   1. stable LSD radix sort of (key, position) pairs (CUB device primitive) -> perm
//   2. run-length encode the sorted keys -> unique keys, counts -> exclusive scan -> segment offsets
//   3. one warp per segment accumulates coef * source-row in registers and writes the gradient row once.
// Deterministic: the order inside a segment is the stable (position) order.
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace dmt {

__global__ void iota_kernel(int32_t* p, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (int32_t)i;
}

// temp layout: [keys_out n u32][vals_in n i32][counts n i32][cub temp ...]
static size_t cub_temp_bytes(int64_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n, 0, 32);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                       (int32_t*)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n + 1);
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int64_t sort_segments_temp_bytes(int64_t n) {
    if (n < 1) n = 1;
    return (int64_t)(3 * align256((size_t)(n + 1) * 4) + align256(cub_temp_bytes(n)) + 256);
}

__global__ void close_offsets_kernel(int32_t* seg_off, const int32_t* n_seg, int32_t n) {
    // ExclusiveSum over counts[0..n_seg) padded with zeros leaves seg_off[n_seg] = n already; this only pins the
    // terminator in case the padding was not zero.
    if (threadIdx.x == 0 && blockIdx.x == 0) seg_off[n_seg[0]] = n;
}

int sort_segments(const uint32_t* keys, int64_t n, int key_bits, int32_t* perm, int32_t* seg_key, int32_t* seg_off,
                  int32_t* n_seg, void* temp, int64_t temp_bytes, cudaStream_t st) {
    if (n <= 0) {
        DMT_CUDA(cudaMemsetAsync(n_seg, 0, sizeof(int32_t), st));
        DMT_CUDA(cudaMemsetAsync(seg_off, 0, sizeof(int32_t), st));
        return 0;
    }
    if (temp_bytes < sort_segments_temp_bytes(n)) {
        set_error("sort_segments: temp buffer too small");
        return DMT_E_ARG;
    }
    char* base = reinterpret_cast<char*>(temp);
    size_t slot = align256((size_t)(n + 1) * 4);
    uint32_t* keys_out = reinterpret_cast<uint32_t*>(base);
    int32_t* vals_in = reinterpret_cast<int32_t*>(base + slot);
    int32_t* counts = reinterpret_cast<int32_t*>(base + 2 * slot);
    void* cub_temp = base + 3 * slot;
    size_t cub_bytes = cub_temp_bytes(n);
    int blocks = (int)((n + 1023) / 1024);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    iota_kernel<<<blocks, 256, 0, st>>>(vals_in, n);
    DMT_LAUNCH_CHECK();
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    DMT_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys, keys_out, vals_in, perm, (int)n, 0, key_bits,
                                             st));
    DMT_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n + 1) * 4, st));
    DMT_CUDA(cub::DeviceRunLengthEncode::Encode(cub_temp, cub_bytes, keys_out, reinterpret_cast<uint32_t*>(seg_key),
                                                counts, n_seg, (int)n, st));
    DMT_CUDA(cub::DeviceScan::ExclusiveSum(cub_temp, cub_bytes, counts, seg_off, (int)n + 1, st));
    close_offsets_kernel<<<1, 32, 0, st>>>(seg_off, n_seg, (int32_t)n);
    DMT_LAUNCH_CHECK();
    return 0;
}

// One warp per segment; VEC float4 slices per lane (width = VEC*128).
template <int VEC>
__global__ void __launch_bounds__(256) segment_reduce_rows_kernel(const int32_t* __restrict__ perm,
                                                                  const int32_t* __restrict__ seg_key,
                                                                  const int32_t* __restrict__ seg_off, SegRef sr,
                                                                  const float* __restrict__ coef,
                                                                  const int32_t* __restrict__ src_row,
                                                                  const float* __restrict__ src,
                                                                  float* __restrict__ grad,
                                                                  float* __restrict__ bias_grad,
                                                                  const int32_t* __restrict__ active) {
    constexpr int W = VEC * 128;
    int64_t s_lo = sr.lo, s_hi = sr.hi;
    uint32_t key_base = 0;
    if (sr.batch_seg_off != nullptr) {
        if (active != nullptr && active[sr.b] == 0) return;
        s_lo = sr.batch_seg_off[sr.b];
        s_hi = sr.batch_seg_off[sr.b + 1];
        key_base = (uint32_t)sr.b * (uint32_t)sr.key_base_stride;
    } else if (sr.n_seg_dev != nullptr) {
        s_hi = sr.n_seg_dev[0];
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * 8;
    for (int64_t s = s_lo + warp; s < s_hi; s += n_warps) {
        const int e0 = seg_off[s], e1 = seg_off[s + 1];
        const int row_out = (int)((uint32_t)seg_key[s] - key_base);
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float bsum = 0.f;
        for (int eb = e0; eb < e1; eb += 32) {
            int e = eb + lane;
            float c_l = 0.f;
            int r_l = 0;
            if (e < e1) {
                int id = perm[e];
                c_l = coef[id];
                r_l = src_row[id];
            }
            bsum += c_l;
            int cnt = min(32, e1 - eb);
            for (int i = 0; i < cnt; ++i) {
                float c = __shfl_sync(0xffffffffu, c_l, i);
                int r = __shfl_sync(0xffffffffu, r_l, i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float4 x = ld4(src + (int64_t)r * W + v * 128 + lane * 4);
                    acc[v].x = fmaf(c, x.x, acc[v].x);
                    acc[v].y = fmaf(c, x.y, acc[v].y);
                    acc[v].z = fmaf(c, x.z, acc[v].z);
                    acc[v].w = fmaf(c, x.w, acc[v].w);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) st4(grad + (int64_t)row_out * W + v * 128 + lane * 4, acc[v]);
        if (bias_grad != nullptr) {
            bsum = warp_sum(bsum);
            if (lane == 0) bias_grad[row_out] = bsum;
        }
    }
}

int launch_segment_reduce_rows(const int32_t* perm, const int32_t* seg_key, const int32_t* seg_off, SegRef sr,
                               int64_t n_seg_max, const float* coef, const int32_t* src_row, const float* src,
                               int width, float* grad, float* bias_grad, const int32_t* active, cudaStream_t st) {
    if (n_seg_max <= 0) return 0;
    int64_t blocks = (n_seg_max + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define DMT_SEG(V)                                                                                            \
    segment_reduce_rows_kernel<V><<<(int)blocks, 256, 0, st>>>(perm, seg_key, seg_off, sr, coef, src_row, src, \
                                                              grad, bias_grad, active)
    if (width == 128) DMT_SEG(1);
    else if (width == 256) DMT_SEG(2);
    else if (width == 384) DMT_SEG(3);
    else if (width == 512) DMT_SEG(4);
    else {
        set_error("segment_reduce_rows: width must be 128, 256, 384 or 512");
        return DMT_E_ARG;
    }
#undef DMT_SEG
    DMT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- load-balanced variant used by the engine
// A segment (one (batch, column) run of the sorted entries) is cut into chunks of <= kSegChunk entries; one warp
// reduces one chunk with 4 source rows in flight. Single-chunk segments write their gradient row directly;
// multi-chunk segments (popular columns: up to one entry per batch row) write partial rows that a second small
// kernel adds in chunk order -> still deterministic, and the critical path is one chunk, not the longest segment.
template <int VEC>
__device__ __forceinline__ void segment_chunks_body(ChunkedSegs cs, const float* __restrict__ coef,
                                                             const float* __restrict__ src, float* __restrict__ grad,
                                                             float* __restrict__ bias_grad,
                                                             const int32_t* __restrict__ active) {
    constexpr int W = VEC * 128;
    if (active != nullptr && active[cs.b] == 0) return;
    const int lane = threadIdx.x & 31;
    const int c_lo = cs.batch_chunk_off[cs.b], c_hi = cs.batch_chunk_off[cs.b + 1];
    const uint32_t key_base = (uint32_t)cs.b * (uint32_t)cs.n_cols;
    const int warp = blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = gridDim.x * 8;
    for (int c = c_lo + warp; c < c_hi; c += n_warps) {
        const int s = cs.chunk_seg[c];
        const int k = c - cs.seg_chunk_off[s];
        const int n_ch = cs.seg_chunk_off[s + 1] - cs.seg_chunk_off[s];
        const int e0 = cs.seg_off[s] + k * kSegChunk;
        const int e1 = min(cs.seg_off[s + 1], e0 + kSegChunk);
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float bsum = 0.f;
        for (int eb = e0; eb < e1; eb += 32) {
            int e = eb + lane;
            float c_l = 0.f;
            int r_l = 0;
            if (e < e1) {
                int id = cs.perm[e];
                c_l = coef[id];
                r_l = cs.ent_row[id];
            }
            bsum += c_l;
            const int cnt = min(32, e1 - eb);
            int i = 0;
            for (; i + 4 <= cnt; i += 4) {  // 4 independent 1 KB source rows in flight per warp
                float cc[4];
                int rr[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    cc[q] = __shfl_sync(0xffffffffu, c_l, i + q);
                    rr[q] = __shfl_sync(0xffffffffu, r_l, i + q);
                }
                float4 x[4][VEC];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[q][v] = ld4(src + (int64_t)rr[q] * W + v * 128 + lane * 4);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        acc[v].x = fmaf(cc[q], x[q][v].x, acc[v].x);
                        acc[v].y = fmaf(cc[q], x[q][v].y, acc[v].y);
                        acc[v].z = fmaf(cc[q], x[q][v].z, acc[v].z);
                        acc[v].w = fmaf(cc[q], x[q][v].w, acc[v].w);
                    }
            }
            for (; i < cnt; ++i) {
                const float c1 = __shfl_sync(0xffffffffu, c_l, i);
                const int r1 = __shfl_sync(0xffffffffu, r_l, i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float4 x = ld4(src + (int64_t)r1 * W + v * 128 + lane * 4);
                    acc[v].x = fmaf(c1, x.x, acc[v].x);
                    acc[v].y = fmaf(c1, x.y, acc[v].y);
                    acc[v].z = fmaf(c1, x.z, acc[v].z);
                    acc[v].w = fmaf(c1, x.w, acc[v].w);
                }
            }
        }
        bsum = warp_sum(bsum);
        if (n_ch == 1) {
            const int row_out = (int)((uint32_t)cs.seg_key[s] - key_base);
#pragma unroll
            for (int v = 0; v < VEC; ++v) st4(grad + (int64_t)row_out * W + v * 128 + lane * 4, acc[v]);
            if (bias_grad != nullptr && lane == 0) bias_grad[row_out] = bsum;
        } else {
            const int64_t slot = c - c_lo;
#pragma unroll
            for (int v = 0; v < VEC; ++v) st4(cs.part + slot * W + v * 128 + lane * 4, acc[v]);
            if (lane == 0) cs.part_bias[slot] = bsum;
        }
    }
}

template <int VEC>
__device__ __forceinline__ void segment_finish_body(ChunkedSegs cs, float* __restrict__ grad,
                                                             float* __restrict__ bias_grad,
                                                             const int32_t* __restrict__ active) {
    constexpr int W = VEC * 128;
    if (active != nullptr && active[cs.b] == 0) return;
    const int lane = threadIdx.x & 31;
    const int s_lo = cs.batch_seg_off[cs.b], s_hi = cs.batch_seg_off[cs.b + 1];
    const int c_base = cs.batch_chunk_off[cs.b];
    const uint32_t key_base = (uint32_t)cs.b * (uint32_t)cs.n_cols;
    const int warp = blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = gridDim.x * 8;
    for (int s = s_lo + warp; s < s_hi; s += n_warps) {
        const int c0 = cs.seg_chunk_off[s], c1 = cs.seg_chunk_off[s + 1];
        if (c1 - c0 <= 1) continue;
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float bsum = 0.f;
        for (int c = c0; c < c1; ++c) {
            const int64_t slot = c - c_base;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float4 x = ld4(cs.part + slot * W + v * 128 + lane * 4);
                acc[v].x += x.x; acc[v].y += x.y; acc[v].z += x.z; acc[v].w += x.w;
            }
            bsum += cs.part_bias[slot];
        }
        const int row_out = (int)((uint32_t)cs.seg_key[s] - key_base);
#pragma unroll
        for (int v = 0; v < VEC; ++v) st4(grad + (int64_t)row_out * W + v * 128 + lane * 4, acc[v]);
        if (bias_grad != nullptr && lane == 0) bias_grad[row_out] = bsum;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) segment_chunks_kernel(ChunkedSegs cs, const float* coef, const float* src,
                                                             float* grad, float* bias_grad, const int32_t* active) {
    segment_chunks_body<VEC>(cs, coef, src, grad, bias_grad, active);
}
template <int VEC>
__global__ void __launch_bounds__(256) segment_finish_kernel(ChunkedSegs cs, float* grad, float* bias_grad,
                                                             const int32_t* active) {
    segment_finish_body<VEC>(cs, grad, bias_grad, active);
}
// side 0: dW4/db4 from (gbuf, A3); side 1: dW1t from (data values, dZ1)
template <int VEC, bool FINISH>
__global__ void __launch_bounds__(256) segment_group(const OrgDev* __restrict__ orgs, int b, int side) {
    const OrgDev& o = orgs[blockIdx.z];
    ChunkedSegs cs = side == 0 ? o.seg_t : o.seg_d;
    cs.b = b;
    const float* coef = side == 0 ? o.gbuf : o.dval_ord;
    const float* src = side == 0 ? o.a3 : o.dz1;
    float* grad = o.G + (side == 0 ? o.oW4 : o.oW1);
    float* bias = side == 0 ? o.G + o.ob4 : nullptr;
    if (FINISH) segment_finish_body<VEC>(cs, grad, bias, o.active);
    else segment_chunks_body<VEC>(cs, coef, src, grad, bias, o.active);
}

int launch_group_segments(const OrgDev* orgs, int G, int b, int side, int n_cols_max, int H1, cudaStream_t st) {
    int per_org = (kNumSMs * 8 + G - 1) / G;
    int want = (n_cols_max * 2 + 7) / 8;
    if (per_org > want) per_org = want;
    if (per_org < 4) per_org = 4;
    int fin = (n_cols_max + 7) / 8;
    if (fin > per_org) fin = per_org;
    if (fin < 1) fin = 1;
#define DMT_SEGG(V)                                                                        \
    do {                                                                                   \
        segment_group<V, false><<<dim3(per_org, 1, G), 256, 0, st>>>(orgs, b, side);       \
        DMT_LAUNCH_CHECK();                                                                \
        segment_group<V, true><<<dim3(fin, 1, G), 256, 0, st>>>(orgs, b, side);            \
        DMT_LAUNCH_CHECK();                                                                \
    } while (0)
    if (H1 == 128) DMT_SEGG(1);
    else if (H1 == 256) DMT_SEGG(2);
    else if (H1 == 384) DMT_SEGG(3);
    else if (H1 == 512) DMT_SEGG(4);
    else { set_error("segment width must be 128, 256, 384 or 512"); return DMT_E_ARG; }
#undef DMT_SEGG
    return 0;
}

int launch_segment_chunks(ChunkedSegs cs, int n_chunk_max, int n_seg_max, const float* coef, const float* src,
                          int width, float* grad, float* bias_grad, const int32_t* active, cudaStream_t st) {
    int blocks = (n_chunk_max + 7) / 8;
    // two per SM also with many organizations per GPU: capping at 148 / 74 measured 206.9 / 225.0 ms per ML1M round
    // against 204.3 (unlike the register-heavy decoder chunk kernel, dmt_org_set_decoder_blocks)
    if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
    if (blocks < 1) blocks = 1;
    int fblocks = (n_seg_max + 7) / 8;
    if (fblocks > kNumSMs) fblocks = kNumSMs;
    if (fblocks < 1) fblocks = 1;
#define DMT_SEGC(V)                                                                                   \
    do {                                                                                              \
        segment_chunks_kernel<V><<<blocks, 256, 0, st>>>(cs, coef, src, grad, bias_grad, active);     \
        DMT_LAUNCH_CHECK();                                                                           \
        segment_finish_kernel<V><<<fblocks, 256, 0, st>>>(cs, grad, bias_grad, active);               \
        DMT_LAUNCH_CHECK();                                                                           \
    } while (0)
    if (width == 128) DMT_SEGC(1);
    else if (width == 256) DMT_SEGC(2);
    else if (width == 384) DMT_SEGC(3);
    else if (width == 512) DMT_SEGC(4);
    else {
        set_error("segment_chunks: width must be 128, 256, 384 or 512");
        return DMT_E_ARG;
    }
#undef DMT_SEGC
    return 0;
}

// chunk tables: n_ch[s] = ceil(len_s / kSegChunk) for s < n_seg (0 beyond), then chunk -> segment map
__global__ void seg_chunk_count_kernel(const int32_t* __restrict__ seg_off, const int32_t* __restrict__ n_seg,
                                       int64_t cap, int32_t* __restrict__ n_ch) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > cap) return;
    int ns = n_seg[0];
    n_ch[s] = (s < ns) ? (seg_off[s + 1] - seg_off[s] + kSegChunk - 1) / kSegChunk : 0;
}

__global__ void seg_chunk_fill_kernel(const int32_t* __restrict__ seg_chunk_off, const int32_t* __restrict__ n_seg,
                                      int32_t* __restrict__ chunk_seg) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg[0]) return;
    for (int c = seg_chunk_off[s]; c < seg_chunk_off[s + 1]; ++c) chunk_seg[c] = s;
}

int build_seg_chunks(const int32_t* seg_off, const int32_t* n_seg, int64_t cap, int32_t* n_ch, int32_t* seg_chunk_off,
                     int32_t* chunk_seg, void* temp, int64_t temp_bytes, cudaStream_t st) {
    if (cap < 0) cap = 0;
    seg_chunk_count_kernel<<<(int)((cap + 1 + 255) / 256), 256, 0, st>>>(seg_off, n_seg, cap, n_ch);
    DMT_LAUNCH_CHECK();
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, n_ch, seg_chunk_off, (int)cap + 1);
    if ((int64_t)bytes > temp_bytes) {
        set_error("build_seg_chunks: temp too small");
        return DMT_E_STATE;
    }
    DMT_CUDA(cub::DeviceScan::ExclusiveSum(temp, bytes, n_ch, seg_chunk_off, (int)cap + 1, st));
    if (cap > 0) {
        seg_chunk_fill_kernel<<<(int)((cap + 255) / 256), 256, 0, st>>>(seg_chunk_off, n_seg, chunk_seg);
        DMT_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int64_t dmt_sort_segments_temp_bytes(int64_t n) { return sort_segments_temp_bytes(n); }

int dmt_sort_segments(const int32_t* keys, int64_t n, int32_t key_bound, int32_t* perm, int32_t* seg_key,
                      int32_t* seg_off, int32_t* n_seg, void* temp, int64_t temp_bytes, void* stream) {
    DMT_REQUIRE(n >= 0 && n < (1LL << 31) - 1 && key_bound > 0, "dmt_sort_segments: bad argument");
    int bits = 1;
    while (bits < 32 && (1LL << bits) < (int64_t)key_bound) ++bits;
    return sort_segments(reinterpret_cast<const uint32_t*>(keys), n, bits, perm, seg_key, seg_off, n_seg, temp,
                         temp_bytes, as_stream(stream));
}

int dmt_segment_reduce_rows(const int32_t* perm, const int32_t* seg_key, const int32_t* seg_off, const int32_t* n_seg,
                            int64_t n_seg_max, const float* coef, const int32_t* src_row, const float* src, int width,
                            float* grad, float* bias_grad, void* stream) {
    DMT_REQUIRE(n_seg_max >= 0, "dmt_segment_reduce_rows: bad argument");
    SegRef sr{nullptr, n_seg, 0, 0, n_seg_max, 0};
    return launch_segment_reduce_rows(perm, seg_key, seg_off, sr, n_seg_max, coef, src_row, src, width, grad,
                                      bias_grad, nullptr, as_stream(stream));
}

}  // extern "C"
