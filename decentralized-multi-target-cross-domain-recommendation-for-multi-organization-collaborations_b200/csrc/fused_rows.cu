// Row-local forward / backward kernels of the fused AAE step with the two small weight matrices STREAMED THROUGH SHARED
// MEMORY by the bulk-copy engine (reference: the Linear(256->128) / Linear(128->256) pair of src/models/ae.py:14-19,
// 44-45 and its backward inside Organization.train, src/organization.py:149-162).
//
// The first version of these kernels (fused.cu: stream_matvec) had every thread pull its weight column from L2 with
// strided scalar loads, 64 values in flight per thread: ~100 registers and ~19 us per launch, all of it L2 latency.
// Here one elected thread issues eight 32 KB cp.async.bulk copies per CTA (W2t then W3t for the forward pass, W3 then
// W2 for the backward pass: each matrix is four contiguous 32 KB slabs along the reduction index) into a three-slot
// shared-memory ring completed on mbarriers; the first three slabs are requested BEFORE the encoder gather / the dZ3
// staging, so the weights arrive while the CTA is still busy with its own rows. All 256 threads then run a small
// register-tiled FFMA product against the slab (conflict-free 128-bit shared loads of the weights, warp-broadcast
// loads of the activations). fp32 FFMA throughout (parity bar 1e-5 on the loss).
//
// Row tile R per CTA is chosen by batch size (8 for 500-row batches; 2 for 100-row batches, e.g. Douban-shape), so small
// batches still fill the machine. CTAs of R >= 4 rows run 512 threads: ncu showed the products bound by the latency of
// their shared-memory operand loads at 8 warps per SM, and the encoder bound by the longest row of the tile when one
// warp walks one row, so every row's data entries are split over NT / 32 / R warps (partials summed in warp order).
#include "bulk.cuh"
#include "kernels.cuh"

namespace dmt {

namespace {

constexpr int H1c = 256, H2c = 128;
constexpr int kSlabFloats = 8192;            // 32 KB
constexpr uint32_t kSlabBytes = kSlabFloats * 4;
constexpr int kStages = 3;
constexpr int kSlabs = 8;                    // 2 matrices x 128 KB

__device__ __forceinline__ void fma4(float4& acc, float c, const float4& x) {
    acc.x = fmaf(c, x.x, acc.x);
    acc.y = fmaf(c, x.y, acc.y);
    acc.z = fmaf(c, x.z, acc.z);
    acc.w = fmaf(c, x.w, acc.w);
}

// The weight stream of one CTA: slab i of 8 lives in ring slot i % 3.
struct WeightStream {
    const float* ring;   // generic address of slot 0
    uint32_t ring_s;     // shared-window address of slot 0
    uint32_t bars_s;     // shared-window address of the three mbarriers
    const float* first;  // slabs 0..3
    const float* second; // slabs 4..7

    __device__ __forceinline__ void issue(int i) const {  // one thread
        const int slot = i % kStages;
        const float* src = (i < 4 ? first : second) + (int64_t)(i & 3) * kSlabFloats;
        const uint32_t bar = bars_s + 8u * (uint32_t)slot;
        bulk::mbar_expect_tx(bar, kSlabBytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring_s + kSlabBytes * (uint32_t)slot),
                     "l"(src), "r"(kSlabBytes), "r"(bar)
                     : "memory");
    }
    __device__ __forceinline__ const float* wait(int i) const {  // every thread
        bulk::mbar_wait(bars_s + 8u * (uint32_t)(i % kStages), (uint32_t)(i / kStages) & 1u);
        return ring + (i % kStages) * kSlabFloats;
    }
    // every thread, after its last read of slab i: the slot is handed back and refilled with slab i + 3
    __device__ __forceinline__ void release(int i) const {
        __syncthreads();
        if (threadIdx.x == 0 && i + kStages < kSlabs) issue(i + kStages);
    }
};

// `early` slabs are requested at once (3 = the whole ring; 2 leaves the last slot to the caller as scratch until it
// calls late_issue after a block barrier)
__device__ __forceinline__ WeightStream stream_setup(float* ring, uint64_t* bars, const float* first,
                                                     const float* second, int early) {
    WeightStream ws{ring, bulk::smem_u32(ring), bulk::smem_u32(bars), first, second};
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) bulk::mbar_init(ws.bars_s + 8u * s, 1);
        bulk::fence_barrier_init();
        for (int i = 0; i < early; ++i) ws.issue(i);
    }
    return ws;
}
__device__ __forceinline__ void late_issue(const WeightStream& ws, int i) {  // after a block barrier
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes to the slot -> bulk-copy writes
        ws.issue(i);
    }
}

// Thread tiling of an [R x N] output over NT threads: CP consecutive columns x RP rows per thread, a warp covering
// 32 consecutive column groups of one row group (weights: conflict-free vector loads; activations: broadcasts).
template <int R, int N, int NT>
struct Tile {
    static constexpr int CP = (R * N / NT) >= 4 ? 4 : (R * N / NT);
    static constexpr int RP = R * N / (NT * CP);
    static constexpr int NCG = N / CP;
    static_assert(CP >= 1 && RP >= 1 && NCG * (R / RP) == NT && NCG % 32 == 0, "tile");
};
constexpr int threads_for(int R) { return R >= 4 ? 512 : 256; }

// acc[rp][cp] += sum_{k in slab} act[row(rp)][k0 + k] * W[k][col0 + cp] for the KC reduction rows of one slab
template <int R, int N, int NT, int KC, int LDA>
__device__ __forceinline__ void slab_product(float (&acc)[Tile<R, N, NT>::RP][Tile<R, N, NT>::CP],
                                             const float* __restrict__ Ws, const float (*act)[LDA], int k0, int row0,
                                             int col0) {
    using T = Tile<R, N, NT>;
#pragma unroll 2
    for (int kk = 0; kk < KC; kk += 4) {
        float4 a[T::RP];
#pragma unroll
        for (int i = 0; i < T::RP; ++i) a[i] = *reinterpret_cast<const float4*>(&act[row0 + i][k0 + kk]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float w[T::CP];
            const float* wp = Ws + (kk + j) * N + col0;
            if constexpr (T::CP == 4) {
                const float4 v = *reinterpret_cast<const float4*>(wp);
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else if constexpr (T::CP == 2) {
                const float2 v = *reinterpret_cast<const float2*>(wp);
                w[0] = v.x; w[1] = v.y;
            } else {
                w[0] = wp[0];
            }
#pragma unroll
            for (int i = 0; i < T::RP; ++i) {
                const float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
#pragma unroll
                for (int c = 0; c < T::CP; ++c) acc[i][c] = fmaf(av, w[c], acc[i][c]);
            }
        }
    }
}

// out[R x N] = act[R x K] * W[K x N] with W arriving as slabs s0 .. s0+3 of the stream (K * N = 4 slabs)
template <int R, int N, int NT, int K, int LDA>
__device__ __forceinline__ void stream_product(const WeightStream& ws, int s0,
                                               float (&acc)[Tile<R, N, NT>::RP][Tile<R, N, NT>::CP],
                                               const float (*act)[LDA], int row0, int col0) {
    constexpr int KC = kSlabFloats / N;  // reduction rows per slab
    static_assert(KC * 4 == K, "four slabs per matrix");
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const float* Ws = ws.wait(s0 + i);
        slab_product<R, N, NT, KC, LDA>(acc, Ws, act, i * KC, row0, col0);
        ws.release(s0 + i);
    }
}

// Shared-memory carve-up (dynamic): ring | mbarriers | activation tiles. The encoder's per-warp partial rows
// (NT / 32 x 1 KB <= 16 KB) live in ring slot 2 until the weights need it.
template <int R>
struct Smem {
    static constexpr int kRing = kStages * kSlabFloats;               // floats
    static constexpr int kWide = R * H1c;                             // [R][256]
    static constexpr int kNarrow = R * H2c;                           // [R][128]
    static constexpr size_t bytes = (size_t)(kRing + kWide + kNarrow) * 4 + 64;
};

// ------------------------------------------------------------------------------------------------ forward rows
template <int R>
__global__ void __launch_bounds__(threads_for(R)) ae_fwd_rows_tma_kernel(FusedFwd p) {
    constexpr int NT = threads_for(R), NW = NT / 32, WPR = NW / R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Smem<R>::kRing);
    float(*a1s)[H1c] = reinterpret_cast<float(*)[H1c]>(reinterpret_cast<float*>(bars) + 16);
    float(*cs)[H2c] = reinterpret_cast<float(*)[H2c]>(&a1s[R][0]);
    float* enc_part = ring + (kStages - 1) * kSlabFloats;  // ring slot 2, free until late_issue below
    static_assert(NW * H1c <= kSlabFloats, "encoder partials fit one ring slot");
    DMT_PDL_ENTRY();
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int m = hi - lo;
    const int r0 = blockIdx.x * R;
    if (r0 >= m) return;
    const WeightStream ws = stream_setup(ring, bars, p.W2t, p.W3t, WPR > 1 ? 2 : 3);  // the weights start travelling
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // ---- encoder: WPR warps per row, each on a contiguous part of the row's data entries, eight 1 KB weight rows of
    //      W1t in flight per warp
    {
        const int r = wid % R, part = wid / R;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        const bool valid = r0 + r < m;
        if (valid) {
            const int u = p.rows[lo + r0 + r];
            const int eA = p.d_indptr[u], eB = p.d_indptr[u + 1];
            const int per = (eB - eA + WPR - 1) / WPR;
            const int e0 = eA + part * per, e1 = min(eB, e0 + per);
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                int c_l = 0;
                float v_l = 0.f;
                if (e < e1) {
                    c_l = p.d_indices[e];
                    v_l = p.d_val[e];
                }
                const int cnt = min(32, e1 - eb);
                for (int i = 0; i < cnt; i += 8) {  // slots past cnt weigh 0
                    float vv[8];
                    float4 w0[8], w1[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int src = min(i + q, cnt - 1);
                        const int col = __shfl_sync(0xffffffffu, c_l, src);
                        const float v = __shfl_sync(0xffffffffu, v_l, src);
                        vv[q] = (i + q < cnt) ? v : 0.f;
                        const float* wr = p.W1t + (int64_t)col * H1c + lane * 4;
                        w0[q] = ld4(wr);
                        w1[q] = ld4(wr + 128);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        fma4(acc0, vv[q], w0[q]);
                        fma4(acc1, vv[q], w1[q]);
                    }
                }
            }
        }
        if constexpr (WPR > 1) {
            st4(enc_part + (part * R + r) * H1c + lane * 4, acc0);
            st4(enc_part + (part * R + r) * H1c + 128 + lane * 4, acc1);
            __syncthreads();
            if (part == 0) {
#pragma unroll
                for (int q = 1; q < WPR; ++q) {  // partials of a row are added in warp order
                    const float4 x0 = ld4(enc_part + (q * R + r) * H1c + lane * 4);
                    const float4 x1 = ld4(enc_part + (q * R + r) * H1c + 128 + lane * 4);
                    acc0.x += x0.x; acc0.y += x0.y; acc0.z += x0.z; acc0.w += x0.w;
                    acc1.x += x1.x; acc1.y += x1.y; acc1.z += x1.z; acc1.w += x1.w;
                }
            }
        }
        if (part == 0) {
            if (valid) {
                const float4 bb0 = ld4(p.b1 + lane * 4), bb1 = ld4(p.b1 + 128 + lane * 4);
                acc0 = make_float4(tanhf(acc0.x + bb0.x), tanhf(acc0.y + bb0.y), tanhf(acc0.z + bb0.z),
                                   tanhf(acc0.w + bb0.w));
                acc1 = make_float4(tanhf(acc1.x + bb1.x), tanhf(acc1.y + bb1.y), tanhf(acc1.z + bb1.z),
                                   tanhf(acc1.w + bb1.w));
                if (p.a1 != nullptr) {
                    st4(p.a1 + (int64_t)(r0 + r) * H1c + lane * 4, acc0);
                    st4(p.a1 + (int64_t)(r0 + r) * H1c + 128 + lane * 4, acc1);
                }
            }
            st4(&a1s[r][lane * 4], acc0);
            st4(&a1s[r][128 + lane * 4], acc1);
        }
    }
    __syncthreads();  // a1s complete; also orders thread 0's barrier initialisation before everybody's first wait
    if constexpr (WPR > 1) late_issue(ws, kStages - 1);  // the partials are consumed: slot 2 joins the weight ring
    // ---- Linear(256 -> 128) + tanh (+ dropout)
    {
        using T = Tile<R, H2c, NT>;
        const int col0 = (t % T::NCG) * T::CP, row0 = (t / T::NCG) * T::RP;
        float acc[T::RP][T::CP];
#pragma unroll
        for (int i = 0; i < T::RP; ++i)
#pragma unroll
            for (int c = 0; c < T::CP; ++c) acc[i][c] = p.b2[col0 + c];
        stream_product<R, H2c, NT, H1c, H1c>(ws, 0, acc, a1s, row0, col0);
#pragma unroll
        for (int i = 0; i < T::RP; ++i) {
            const int row = row0 + i;
#pragma unroll
            for (int c = 0; c < T::CP; ++c) {
                const int n = col0 + c;
                float v = tanhf(acc[i][c]);
                if (r0 + row < m) {
                    if (p.a2 != nullptr) p.a2[(int64_t)(r0 + row) * H2c + n] = v;
                    if (p.drop.enabled) v *= dropout_factor(p.drop, r0 + row, n, H2c);
                    if (p.c != nullptr) p.c[(int64_t)(r0 + row) * H2c + n] = v;
                } else {
                    v = 0.f;
                }
                cs[row][n] = v;
            }
        }
    }
    __syncthreads();
    // ---- Linear(128 -> 256) + tanh
    {
        using T = Tile<R, H1c, NT>;
        const int col0 = (t % T::NCG) * T::CP, row0 = (t / T::NCG) * T::RP;
        float acc[T::RP][T::CP];
#pragma unroll
        for (int i = 0; i < T::RP; ++i)
#pragma unroll
            for (int c = 0; c < T::CP; ++c) acc[i][c] = p.b3[col0 + c];
        stream_product<R, H1c, NT, H2c, H2c>(ws, 4, acc, cs, row0, col0);
#pragma unroll
        for (int i = 0; i < T::RP; ++i)
            if (r0 + row0 + i < m) {
#pragma unroll
                for (int c = 0; c < T::CP; ++c)
                    p.a3[(int64_t)(r0 + row0 + i) * H1c + col0 + c] = tanhf(acc[i][c]);
            }
    }
}

// ------------------------------------------------------------------------------------------------ backward rows
// dZ3 -> dZ2 -> dZ1 for a tile of R rows; per-CTA column sums for db3 | db2 | db1 at part_db[cta * 768 + {0, 256, 512}]
template <int R>
__global__ void __launch_bounds__(threads_for(R)) ae_bwd_rows_tma_kernel(FusedBwd p) {
    constexpr int NT = threads_for(R);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Smem<R>::kRing);
    float(*wide)[H1c] = reinterpret_cast<float(*)[H1c]>(reinterpret_cast<float*>(bars) + 16);  // dZ3, later dZ1
    float(*d2s)[H2c] = reinterpret_cast<float(*)[H2c]>(&wide[R][0]);
    DMT_PDL_ENTRY();
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int m = hi - lo;
    const int r0 = blockIdx.x * R;
    if (r0 >= m) return;
    const WeightStream ws = stream_setup(ring, bars, p.W3, p.W2, 3);
    const int t = threadIdx.x;
    float* part = p.part_db + (int64_t)blockIdx.x * kDbPartStride;
    // dZ3 rows of the tile (rows without targets received no decoder chunk: their dZ3 is zero)
    for (int idx = t; idx < R * H1c; idx += NT) {
        const int i = idx / H1c, col = idx % H1c;
        float v = 0.f;
        if (r0 + i < m && p.t_len[lo + r0 + i] > 0) v = p.dz3[(int64_t)(r0 + i) * H1c + col];
        wide[i][col] = v;
    }
    __syncthreads();
    if (t < H1c) {  // db3 partial
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i) s += wide[i][t];
        part[t] = s;
    }
    // dZ2 = (dZ3 W3) * dropout * (1 - a2^2); W3 is [256 x 128] row-major: four slabs of 64 reduction rows
    {
        using T = Tile<R, H2c, NT>;
        const int col0 = (t % T::NCG) * T::CP, row0 = (t / T::NCG) * T::RP;
        float acc[T::RP][T::CP];
#pragma unroll
        for (int i = 0; i < T::RP; ++i)
#pragma unroll
            for (int c = 0; c < T::CP; ++c) acc[i][c] = 0.f;
        stream_product<R, H2c, NT, H1c, H1c>(ws, 0, acc, wide, row0, col0);
#pragma unroll
        for (int i = 0; i < T::RP; ++i) {
            const int row = row0 + i;
#pragma unroll
            for (int c = 0; c < T::CP; ++c) {
                const int k = col0 + c;
                float v = 0.f;
                if (r0 + row < m) {
                    v = acc[i][c];
                    if (p.drop.enabled) v *= dropout_factor(p.drop, r0 + row, k, H2c);
                    const float av = p.a2[(int64_t)(r0 + row) * H2c + k];
                    v *= 1.f - av * av;
                    p.dz2[(int64_t)(r0 + row) * H2c + k] = v;
                }
                d2s[row][k] = v;
            }
        }
    }
    __syncthreads();
    if (t < H2c) {  // db2 partial
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i) s += d2s[i][t];
        part[H1c + t] = s;
        part[H1c + H2c + t] = 0.f;  // (the register-streamed kernel keeps two half-tile partials here)
    }
    // dZ1 = (dZ2 W2) * (1 - a1^2); W2 is [128 x 256] row-major: four slabs of 32 reduction rows
    {
        using T = Tile<R, H1c, NT>;
        const int col0 = (t % T::NCG) * T::CP, row0 = (t / T::NCG) * T::RP;
        float acc[T::RP][T::CP];
#pragma unroll
        for (int i = 0; i < T::RP; ++i)
#pragma unroll
            for (int c = 0; c < T::CP; ++c) acc[i][c] = 0.f;
        stream_product<R, H1c, NT, H2c, H2c>(ws, 4, acc, d2s, row0, col0);
        // (the last release() of the stream was a block barrier: every thread is done with the dZ3 tile in `wide`,
        //  and the db3 sums above were taken before the first one)
#pragma unroll
        for (int i = 0; i < T::RP; ++i) {
            const int row = row0 + i;
#pragma unroll
            for (int c = 0; c < T::CP; ++c) {
                const int k = col0 + c;
                float v = 0.f;
                if (r0 + row < m) {
                    const float av = p.a1[(int64_t)(r0 + row) * H1c + k];
                    v = acc[i][c] * (1.f - av * av);
                    p.dz1[(int64_t)(r0 + row) * H1c + k] = v;
                }
                wide[row][k] = v;
            }
        }
    }
    __syncthreads();
    if (t < H1c) {  // db1 partial
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i) s += wide[i][t];
        part[2 * H1c + t] = s;
    }
}

template <int R>
int allow_smem() {
    DMT_CUDA(cudaFuncSetAttribute(ae_fwd_rows_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Smem<R>::bytes));
    DMT_CUDA(cudaFuncSetAttribute(ae_bwd_rows_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Smem<R>::bytes));
    return 0;
}

}  // namespace

int fused_rows_per_cta(int batch_rows) { return batch_rows <= 128 ? 2 : (batch_rows <= 256 ? 4 : 8); }

// Opt the kernels into > 48 KB of dynamic shared memory on the CURRENT device (call once per device before the first
// launch, outside stream capture; cheap enough to repeat per organization).
int prepare_fused_rows() {
    int rc;
    if ((rc = allow_smem<2>()) || (rc = allow_smem<4>()) || (rc = allow_smem<8>())) return rc;
    return 0;
}

int launch_fused_fwd_tma(const FusedFwd& p, int n_rows_max, int R, cudaStream_t st, bool pdl) {
    if (n_rows_max <= 0) return 0;
    const int grid = (n_rows_max + R - 1) / R;
    switch (R) {
        case 2:
            DMT_CUDA(launch_k(ae_fwd_rows_tma_kernel<2>, grid, threads_for(2), Smem<2>::bytes, st, pdl, p));
            break;
        case 4:
            DMT_CUDA(launch_k(ae_fwd_rows_tma_kernel<4>, grid, threads_for(4), Smem<4>::bytes, st, pdl, p));
            break;
        default:
            DMT_REQUIRE(R == 8, "launch_fused_fwd_tma: row tile must be 2, 4 or 8");
            DMT_CUDA(launch_k(ae_fwd_rows_tma_kernel<8>, grid, threads_for(8), Smem<8>::bytes, st, pdl, p));
    }
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_fused_bwd_rows_tma(const FusedBwd& p, int n_rows_max, int R, cudaStream_t st, bool pdl) {
    if (n_rows_max <= 0) return 0;
    const int grid = (n_rows_max + R - 1) / R;
    switch (R) {
        case 2:
            DMT_CUDA(launch_k(ae_bwd_rows_tma_kernel<2>, grid, threads_for(2), Smem<2>::bytes, st, pdl, p));
            break;
        case 4:
            DMT_CUDA(launch_k(ae_bwd_rows_tma_kernel<4>, grid, threads_for(4), Smem<4>::bytes, st, pdl, p));
            break;
        default:
            DMT_REQUIRE(R == 8, "launch_fused_bwd_rows_tma: row tile must be 2, 4 or 8");
            DMT_CUDA(launch_k(ae_bwd_rows_tma_kernel<8>, grid, threads_for(8), Smem<8>::bytes, st, pdl, p));
    }
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmt
