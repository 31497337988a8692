// Fused training step of the AAE engine: SIX dependent launches per batch instead of twenty
// (reference: one iteration of the loop in src/organization.py:149-162 around AE.forward, src/models/ae.py:98-157).
//
//   1 ae_fwd_rows      row-local forward: CSR SpMM encoder + tanh -> Linear(256->128) + tanh + dropout ->
//                      Linear(128->256) + tanh for a tile of FR batch rows per CTA; the two small weight matrices are
//                      streamed from L2 in their TRANSPOSED shadow layout (coalesced along the output unit), the
//                      activations of the tile stay in shared memory between the layers.
//   2 ae_dec_chunks    decoder SDDMM + loss + g = dL/do + partial dZ3 per chunk of <= 128 targets of one row; the LAST
//                      chunk of a row to arrive adds the row's partials in chunk order (deterministic) and applies the
//                      tanh derivative -> no finish kernel. g is scattered straight into (batch, column)-sorted order.
//   3 ae_bwd_rows      row-local backward dZ3 -> dZ2 -> dZ1 with per-CTA column-sum partials for db3 / db2 / db1;
//     ae_seg_chunks    on a PARALLEL branch of the step graph (it only has to finish before the norm): dW4 / db4 as a
//                      chunked segmented reduction over the sorted targets (last-arriving chunk of a multi-chunk
//                      segment adds the partial rows in order).
//   4 ae_grad_phase    three block roles: (a) dW3 = dZ3^T C and dW2 = dZ2^T A1 as 64x64 FFMA tiles, split over row
//                      slices, last slice to arrive adds the slices in order, (b) bias-gradient finish, (c) dW1t as a
//                      chunked segmented reduction over the sorted data entries.
//   5 norm_prepare     sum of squares of the flat gradient; the last block derives the clip coefficient, the Adam
//                      bias corrections, the batch loss, and advances the step counter.
//   6 adam             dense Adam(+L2) with the clip folded in; it also refreshes the transposed shadows of W2 / W3.
//
// All kernels read their batch bounds from device memory (BatchRef), so the epoch graph is replayable. fp32 FFMA
// throughout (the parity bar is 1e-5 relative on the loss). Shapes: H1 = 256, H2 = 128 (src/utils.py:166-171).
#include "bulk.cuh"
#include "kernels.cuh"

namespace dmt {

namespace {

constexpr int H1c = 256, H2c = 128;

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma4(float4& acc, float c, const float4& x) {
    acc.x = fmaf(c, x.x, acc.x);
    acc.y = fmaf(c, x.y, acc.y);
    acc.z = fmaf(c, x.z, acc.z);
    acc.w = fmaf(c, x.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, const float4& x) {
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}


// acc[i] += sum_k act[row0 + i][k] * W[k * ld + col] for k < K: the weight column is streamed from L2 in batches of
// PF values per thread, the next batch in flight while the current one is consumed (these row-local kernels run few
// CTAs, so memory-level parallelism has to come from each thread); the activations are broadcast from shared memory.
template <int NR, int K, int PF, int LDA>
__device__ __forceinline__ void stream_matvec(float (&acc)[NR], const float* __restrict__ w, int ld,
                                              const float (*act)[LDA], int row0) {
    static_assert(K % PF == 0 && PF % 4 == 0, "prefetch batch");
    float cur[PF], nxt[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) cur[i] = w[(int64_t)i * ld];
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += PF) {
        if (k0 + PF < K) {
#pragma unroll
            for (int i = 0; i < PF; ++i) nxt[i] = w[(int64_t)(k0 + PF + i) * ld];
        }
#pragma unroll
        for (int q = 0; q < PF / 4; ++q) {
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(&act[row0 + i][k0 + 4 * q]);
                acc[i] = fmaf(a.x, cur[4 * q + 0], acc[i]);
                acc[i] = fmaf(a.y, cur[4 * q + 1], acc[i]);
                acc[i] = fmaf(a.z, cur[4 * q + 2], acc[i]);
                acc[i] = fmaf(a.w, cur[4 * q + 3], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) cur[i] = nxt[i];
    }
}

// ------------------------------------------------------------------------------------------------ 1 forward rows
template <int R>
__device__ __forceinline__ void fwd_rows_body(const FusedFwd& p, int cta) {
    static_assert(R % 2 == 0 && R <= 16, "row tile");
    __shared__ __align__(16) float a1s[R][H1c];
    __shared__ __align__(16) float cs[R][H2c];
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int m = hi - lo;
    const int r0 = cta * R;
    if (r0 >= m) return;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // ---- encoder: every lane owns 2 float4 slices of the 256 hidden units; a row's data entries are split over
    //      WPR = 8 / R warps when the tile has fewer rows than the block has warps (the longest row of the tile bounds
    //      this phase: ~100 entries at ML1M shape), eight 1 KB weight rows in flight per warp
    constexpr int WPR = R >= 8 ? 1 : 8 / R;
    __shared__ __align__(16) float encp[WPR > 1 ? (WPR - 1) * R : 1][H1c];
    for (int rb = 0; rb < (R > 8 ? R : 8); rb += 8) {
        const int r = WPR > 1 ? wid % R : rb + wid;
        const int part = WPR > 1 ? wid / R : 0;
        if (r >= R) break;
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
            if (part > 0) {
                st4(&encp[(part - 1) * R + r][lane * 4], acc0);
                st4(&encp[(part - 1) * R + r][128 + lane * 4], acc1);
            }
            __syncthreads();
            if (part > 0) break;
#pragma unroll
            for (int q = 1; q < WPR; ++q) {  // a row's partials are added in warp order
                add4(acc0, ld4(&encp[(q - 1) * R + r][lane * 4]));
                add4(acc1, ld4(&encp[(q - 1) * R + r][128 + lane * 4]));
            }
        }
        if (valid) {
            const float4 bb0 = ld4(p.b1 + lane * 4), bb1 = ld4(p.b1 + 128 + lane * 4);
            acc0 = make_float4(tanhf(acc0.x + bb0.x), tanhf(acc0.y + bb0.y), tanhf(acc0.z + bb0.z), tanhf(acc0.w + bb0.w));
            acc1 = make_float4(tanhf(acc1.x + bb1.x), tanhf(acc1.y + bb1.y), tanhf(acc1.z + bb1.z), tanhf(acc1.w + bb1.w));
            if (p.a1 != nullptr) {
                st4(p.a1 + (int64_t)(r0 + r) * H1c + lane * 4, acc0);
                st4(p.a1 + (int64_t)(r0 + r) * H1c + 128 + lane * 4, acc1);
            }
        }
        st4(&a1s[r][lane * 4], acc0);
        st4(&a1s[r][128 + lane * 4], acc1);
    }
    __syncthreads();
    // ---- Linear(256 -> 128) + tanh (+ dropout): thread = (output unit n, half of the row tile)
    {
        constexpr int RH = R / 2;
        const int n = t & (H2c - 1), g = t >> 7;
        float acc[RH];
        const float bias = p.b2[n];
#pragma unroll
        for (int i = 0; i < RH; ++i) acc[i] = bias;
        stream_matvec<RH, H1c, 32, H1c>(acc, p.W2t + n, H2c, a1s, g * RH);
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            const int row = g * RH + i;
            float v = tanhf(acc[i]);
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
    __syncthreads();
    // ---- Linear(128 -> 256) + tanh: thread = output unit
    {
        const int n = t;
        float acc[R];
        const float bias = p.b3[n];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = bias;
        stream_matvec<R, H2c, 32, H2c>(acc, p.W3t + n, H1c, cs, 0);
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (r0 + i < m) p.a3[(int64_t)(r0 + i) * H1c + n] = tanhf(acc[i]);
    }
}

// ------------------------------------------------------------------------------------------------ 2 decoder chunks
// One block per chunk of <= kDecChunk targets of one batch row (persistent over the batch's chunks). Per chunk ONE
// metadata load (plan-time int4) replaces the chunk -> row -> CSR pointer chain.
__device__ __forceinline__ void dec_chunks_body(const FusedDec& p) {
    constexpr int VEC = 2, H = H1c;
    constexpr int PER_WARP = kDecChunk / 8;
    constexpr int GROUPS = PER_WARP / 4;
    __shared__ __align__(16) float s_acc[2][8][H];
    __shared__ float s_loss[2][8];
    __shared__ int s_last;
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c_lo = p.batch_chunk_off[p.br.b], c_hi = p.batch_chunk_off[p.br.b + 1];
    const float inv_n = 1.f / (float)p.n_targets[p.br.b];
    int buf = 0;
    for (int c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x, buf ^= 1) {
        // {in-batch row, CSR position of the chunk's first target, its batch-order entry id, packed counts}
        const int4 mt = p.meta[c];
        const int jl = mt.x, e0 = mt.y, out0 = mt.z;
        const int n_tg = mt.w & 0xff;               // 1..128 targets in this chunk
        const int k = (mt.w >> 8) & 0xfff;          // chunk index inside the row
        const int n_ch = (mt.w >> 20) & 0xfff;      // chunks of the row
        const float* a_row = p.A3 + (int64_t)jl * H;
        float4 a[VEC], acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            a[v] = ld4(a_row + v * 128 + lane * 4);
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float loss_acc = 0.f;
        const int eb = e0 + wid * PER_WARP;
        const int cnt = max(0, min(PER_WARP, n_tg - wid * PER_WARP));
        int c_l = 0, pos_l = 0;
        float y_l = 0.f;
        if (lane < cnt) {
            c_l = p.t_indices[eb + lane];
            y_l = p.target[eb + lane];
            pos_l = p.inv_perm[out0 + wid * PER_WARP + lane];
        }
        float o_l = 0.f;
        if (cnt > 0) {
            float4 w[2][4][VEC];
            float bb[2][4];
            auto load_group = [&](int g, int slot) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = __shfl_sync(0xffffffffu, c_l, min(4 * g + q, cnt - 1));
#pragma unroll
                    for (int v = 0; v < VEC; ++v) w[slot][q][v] = ld4(p.W4 + (int64_t)col * H + v * 128 + lane * 4);
                    bb[slot][q] = p.b4[col];
                }
            };
            load_group(0, 0);
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                if (4 * g < cnt) {
                    const int slot = g & 1;
                    if (g + 1 < GROUPS && 4 * (g + 1) < cnt) load_group(g + 1, (g + 1) & 1);
                    float d[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        d[q] = 0.f;
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            d[q] += a[v].x * w[slot][q][v].x + a[v].y * w[slot][q][v].y + a[v].z * w[slot][q][v].z +
                                    a[v].w * w[slot][q][v].w;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int tq = 4 * g + q;
                        const bool valid = tq < cnt;
                        d[q] = warp_sum(d[q]) + bb[slot][q];
                        if (lane == tq && valid) o_l = d[q];
                        const float y = __shfl_sync(0xffffffffu, y_l, min(tq, cnt - 1));
                        const float gq = valid ? loss_grad(DMT_LOSS_MSE, d[q], y) * inv_n : 0.f;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) fma4(acc[v], gq, w[slot][q][v]);
                    }
                }
            }
        }
        if (lane < cnt) {
            p.g_sorted[pos_l] = loss_grad(DMT_LOSS_MSE, o_l, y_l) * inv_n;
            loss_acc = loss_value(DMT_LOSS_MSE, o_l, y_l);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) st4(&s_acc[buf][wid][v * 128 + lane * 4], acc[v]);
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) s_loss[buf][wid] = loss_acc;
        __syncthreads();
        const int h = threadIdx.x;  // 256 threads == H
        float s = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) s += s_acc[buf][w8][h];
        float l = 0.f;
        if (h == 0) {
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) l += s_loss[buf][w8];
        }
        if (n_ch == 1) {
            const float av = a_row[h];
            p.dZ3[(int64_t)jl * H + h] = s * (1.f - av * av);
            if (h == 0) p.loss_rows[jl] = l;
        } else {
            const int64_t slot = c - c_lo;
            p.dz_part[slot * H + h] = s;
            if (h == 0) p.loss_part[slot] = l;
            __threadfence();
            __syncthreads();
            if (h == 0) s_last = (atomicAdd(&p.row_cnt[jl], 1) == n_ch - 1);
            __syncthreads();
            if (s_last) {  // block-uniform: this chunk arrived last, add the row's partials in chunk order
                __threadfence();
                const int64_t first = slot - k;
                float tot = 0.f;
                for (int q = 0; q < n_ch; ++q) tot += __ldcg(p.dz_part + (first + q) * H + h);
                const float av = a_row[h];
                p.dZ3[(int64_t)jl * H + h] = tot * (1.f - av * av);
                if (h == 0) {
                    float lt = 0.f;
                    for (int q = 0; q < n_ch; ++q) lt += __ldcg(p.loss_part + first + q);
                    p.loss_rows[jl] = lt;
                    p.row_cnt[jl] = 0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ 2' decoder, bulk form
// Same contract as dec_chunks_body; the W4 rows of a chunk travel through warp-private shared-memory rings filled by
// the bulk-copy engine (bulk.cuh) instead of through registers. Four warps per block share a chunk's targets evenly;
// ~50 registers per thread and 41 KB of shared memory per block let five blocks live on one SM, each with up to
// 32 KB of weight rows in flight.
constexpr int kBulkWarps = 4;
constexpr int kBulkSlots = 8;

__device__ __forceinline__ void dec_chunks_bulk_body(const FusedDec& p) {
    constexpr int H = H1c;
    __shared__ __align__(128) float ring[kBulkWarps][kBulkSlots][H];
    __shared__ __align__(16) float s_acc[2][kBulkWarps][H];
    __shared__ float s_loss[2][kBulkWarps];
    __shared__ __align__(8) uint64_t bars[kBulkWarps][kBulkSlots];
    __shared__ int s_last;
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c_lo = p.batch_chunk_off[p.br.b], c_hi = p.batch_chunk_off[p.br.b + 1];
    if (c_lo + (int)blockIdx.x >= c_hi) return;
    bulk::WarpRing<kBulkSlots> rg = bulk::ring_setup<kBulkSlots>(&ring[wid][0][0], &bars[wid][0], lane);
    __syncthreads();
    const float inv_n = 1.f / (float)p.n_targets[p.br.b];
    int buf = 0;
    for (int c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x, buf ^= 1) {
        const int4 mt = p.meta[c];
        const int jl = mt.x, e0 = mt.y, out0 = mt.z;
        const int n_tg = mt.w & 0xff;               // 1..128 targets in this chunk
        const int k = (mt.w >> 8) & 0xfff;          // chunk index inside the row
        const int n_ch = (mt.w >> 20) & 0xfff;      // chunks of the row
        const float* a_row = p.A3 + (int64_t)jl * H;
        const int per = (n_tg + kBulkWarps - 1) / kBulkWarps;  // <= 32 targets per warp
        const int off = wid * per;
        const int cnt = max(0, min(per, n_tg - off));
        int c_l = 0, pos_l = 0;
        float y_l = 0.f, b_l = 0.f;
        if (lane < cnt) {
            c_l = p.t_indices[e0 + off + lane];
            y_l = p.target[e0 + off + lane];
            pos_l = p.inv_perm[out0 + off + lane];
            b_l = p.b4[c_l];
        }
        const float4 a0 = ld4(a_row + lane * 4), a1 = ld4(a_row + 128 + lane * 4);
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        float o_l = 0.f;
        bulk::gather_rows(rg, p.W4, c_l, cnt, lane, [&](int t0, int nv, const float4(&w)[4][2]) {
            float d[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = dot4(a0, w[q][0]) + dot4(a1, w[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = warp_sum(d[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q < nv) {
                    const int tq = t0 + q;
                    const float o = d[q] + __shfl_sync(0xffffffffu, b_l, tq);
                    const float y = __shfl_sync(0xffffffffu, y_l, tq);
                    if (lane == tq) o_l = o;
                    const float gq = loss_grad(DMT_LOSS_MSE, o, y) * inv_n;
                    fma4(acc0, gq, w[q][0]);
                    fma4(acc1, gq, w[q][1]);
                }
            }
        });
        float loss_acc = 0.f;
        if (lane < cnt) {
            p.g_sorted[pos_l] = loss_grad(DMT_LOSS_MSE, o_l, y_l) * inv_n;
            loss_acc = loss_value(DMT_LOSS_MSE, o_l, y_l);
        }
        st4(&s_acc[buf][wid][lane * 4], acc0);
        st4(&s_acc[buf][wid][128 + lane * 4], acc1);
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) s_loss[buf][wid] = loss_acc;
        __syncthreads();
        const int h = threadIdx.x * 2;  // 128 threads x 2 hidden units
        float2 s = make_float2(0.f, 0.f);
        float l = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < kBulkWarps; ++w4) {
            const float2 v = *reinterpret_cast<const float2*>(&s_acc[buf][w4][h]);
            s.x += v.x;
            s.y += v.y;
            l += s_loss[buf][w4];
        }
        if (n_ch == 1) {
            const float2 av = *reinterpret_cast<const float2*>(a_row + h);
            *reinterpret_cast<float2*>(p.dZ3 + (int64_t)jl * H + h) =
                make_float2(s.x * (1.f - av.x * av.x), s.y * (1.f - av.y * av.y));
            if (threadIdx.x == 0) p.loss_rows[jl] = l;
        } else {
            const int64_t slot = c - c_lo;
            *reinterpret_cast<float2*>(p.dz_part + slot * H + h) = s;
            if (threadIdx.x == 0) p.loss_part[slot] = l;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(&p.row_cnt[jl], 1) == n_ch - 1);
            __syncthreads();
            if (s_last) {  // block-uniform: this chunk arrived last, add the row's partials in chunk order
                __threadfence();
                const int64_t first = slot - k;
                float2 tot = make_float2(0.f, 0.f);
                for (int q = 0; q < n_ch; ++q) {
                    const float2 v = __ldcg(reinterpret_cast<const float2*>(p.dz_part + (first + q) * H + h));
                    tot.x += v.x;
                    tot.y += v.y;
                }
                const float2 av = *reinterpret_cast<const float2*>(a_row + h);
                *reinterpret_cast<float2*>(p.dZ3 + (int64_t)jl * H + h) =
                    make_float2(tot.x * (1.f - av.x * av.x), tot.y * (1.f - av.y * av.y));
                if (threadIdx.x == 0) {
                    float lt = 0.f;
                    for (int q = 0; q < n_ch; ++q) lt += __ldcg(p.loss_part + first + q);
                    p.loss_rows[jl] = lt;
                    p.row_cnt[jl] = 0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ 2'' decoder, 4 warps
// Same contract as dec_chunks_body with the chunk's latency chain shortened: four warps per block share a chunk's
// targets evenly (up to 32 per warp: eight groups of four rows, so the streaming part of a chunk outweighs its
// prologue), and the NEXT chunk's metadata, column indices, targets and scatter positions are requested while the
// current chunk's rows are still in flight (the meta -> index -> row dependency chain of a chunk costs three L2 round
// trips; two of them now overlap the previous chunk).
__device__ __forceinline__ void dec_chunks4_body(const FusedDec& p) {
    constexpr int H = H1c, NW = 4, PER_WARP = kDecChunk / NW, GROUPS = PER_WARP / 4;
    __shared__ __align__(16) float s_acc[2][NW][H];
    __shared__ float s_loss[2][NW];
    __shared__ int s_last;
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c_lo = p.batch_chunk_off[p.br.b], c_hi = p.batch_chunk_off[p.br.b + 1];
    int c = c_lo + blockIdx.x;
    if (c >= c_hi) return;
    const float inv_n = 1.f / (float)p.n_targets[p.br.b];
    int4 mt = p.meta[c];
    int c_l = 0, pos_l = 0;
    float y_l = 0.f;
    auto lane_inputs = [&](const int4& m4, int& col, float& y, int& pos) {
        const int n_tg = m4.w & 0xff;
        const int per = (n_tg + NW - 1) / NW;
        const int off = wid * per;
        const int cnt = max(0, min(per, n_tg - off));
        col = 0; y = 0.f; pos = 0;
        if (lane < cnt) {
            col = p.t_indices[m4.y + off + lane];
            y = p.target[m4.y + off + lane];
            pos = p.inv_perm[m4.z + off + lane];
        }
    };
    lane_inputs(mt, c_l, y_l, pos_l);
    int buf = 0;
    for (;; buf ^= 1) {
        const int jl = mt.x;
        const int n_tg = mt.w & 0xff;               // 1..128 targets in this chunk
        const int k = (mt.w >> 8) & 0xfff;          // chunk index inside the row
        const int n_ch = (mt.w >> 20) & 0xfff;      // chunks of the row
        const int per = (n_tg + NW - 1) / NW;
        const int cnt = max(0, min(per, n_tg - wid * per));
        const int cn = c + gridDim.x;
        const bool more = cn < c_hi;
        int4 mt_n = mt;
        if (more) mt_n = p.meta[cn];
        int c_n = 0, pos_n = 0;
        float y_n = 0.f;
        const float* a_row = p.A3 + (int64_t)jl * H;
        const float4 a0 = ld4(a_row + lane * 4), a1 = ld4(a_row + 128 + lane * 4);
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        float o_l = 0.f;
        if (cnt > 0) {
            float4 w[2][4][2];
            float bb[2][4];
            auto load_group = [&](int g, int slot) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = __shfl_sync(0xffffffffu, c_l, min(4 * g + q, cnt - 1));
                    const float* wr = p.W4 + (int64_t)col * H + lane * 4;
                    w[slot][q][0] = ld4(wr);
                    w[slot][q][1] = ld4(wr + 128);
                    bb[slot][q] = p.b4[col];
                }
            };
            load_group(0, 0);
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                if (4 * g < cnt) {
                    const int slot = g & 1;
                    if (g + 1 < GROUPS && 4 * (g + 1) < cnt) load_group(g + 1, (g + 1) & 1);
                    if (g == 0 && more) lane_inputs(mt_n, c_n, y_n, pos_n);  // next chunk's inputs travel with ours
                    float d[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] = dot4(a0, w[slot][q][0]) + dot4(a1, w[slot][q][1]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] = warp_sum(d[q]) + bb[slot][q];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int tq = 4 * g + q;
                        const bool valid = tq < cnt;
                        if (lane == tq && valid) o_l = d[q];
                        const float y = __shfl_sync(0xffffffffu, y_l, min(tq, cnt - 1));
                        const float gq = valid ? loss_grad(DMT_LOSS_MSE, d[q], y) * inv_n : 0.f;
                        fma4(acc0, gq, w[slot][q][0]);
                        fma4(acc1, gq, w[slot][q][1]);
                    }
                }
            }
        } else if (more) {
            lane_inputs(mt_n, c_n, y_n, pos_n);
        }
        float loss_acc = 0.f;
        if (lane < cnt) {
            p.g_sorted[pos_l] = loss_grad(DMT_LOSS_MSE, o_l, y_l) * inv_n;
            loss_acc = loss_value(DMT_LOSS_MSE, o_l, y_l);
        }
        st4(&s_acc[buf][wid][lane * 4], acc0);
        st4(&s_acc[buf][wid][128 + lane * 4], acc1);
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) s_loss[buf][wid] = loss_acc;
        __syncthreads();
        const int h = threadIdx.x * 2;  // 128 threads x 2 hidden units
        float2 s = make_float2(0.f, 0.f);
        float l = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < NW; ++w4) {
            const float2 v = *reinterpret_cast<const float2*>(&s_acc[buf][w4][h]);
            s.x += v.x;
            s.y += v.y;
            l += s_loss[buf][w4];
        }
        if (n_ch == 1) {
            const float2 av = *reinterpret_cast<const float2*>(a_row + h);
            *reinterpret_cast<float2*>(p.dZ3 + (int64_t)jl * H + h) =
                make_float2(s.x * (1.f - av.x * av.x), s.y * (1.f - av.y * av.y));
            if (threadIdx.x == 0) p.loss_rows[jl] = l;
        } else {
            const int64_t slot = c - c_lo;
            *reinterpret_cast<float2*>(p.dz_part + slot * H + h) = s;
            if (threadIdx.x == 0) p.loss_part[slot] = l;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(&p.row_cnt[jl], 1) == n_ch - 1);
            __syncthreads();
            if (s_last) {  // block-uniform: this chunk arrived last, add the row's partials in chunk order
                __threadfence();
                const int64_t first = slot - k;
                float2 tot = make_float2(0.f, 0.f);
                for (int q = 0; q < n_ch; ++q) {
                    const float2 v = __ldcg(reinterpret_cast<const float2*>(p.dz_part + (first + q) * H + h));
                    tot.x += v.x;
                    tot.y += v.y;
                }
                const float2 av = *reinterpret_cast<const float2*>(a_row + h);
                *reinterpret_cast<float2*>(p.dZ3 + (int64_t)jl * H + h) =
                    make_float2(tot.x * (1.f - av.x * av.x), tot.y * (1.f - av.y * av.y));
                if (threadIdx.x == 0) {
                    float lt = 0.f;
                    for (int q = 0; q < n_ch; ++q) lt += __ldcg(p.loss_part + first + q);
                    p.loss_rows[jl] = lt;
                    p.row_cnt[jl] = 0;
                }
            }
        }
        if (!more) break;
        c = cn;
        mt = mt_n;
        c_l = c_n;
        y_l = y_n;
        pos_l = pos_n;
    }
}

// ------------------------------------------------------------------------------------------------ 3a backward rows
template <int R>
__device__ __forceinline__ void bwd_rows_body(const FusedBwd& p, int cta) {
    __shared__ __align__(16) float d3s[R][H1c];
    __shared__ __align__(16) float d2s[R][H2c];
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int m = hi - lo;
    const int r0 = cta * R;
    if (r0 >= m) return;
    const int t = threadIdx.x;
    float* part = p.part_db + (int64_t)cta * kDbPartStride;
    // dZ3 rows of the tile (rows without targets received no chunk: their dZ3 is zero)
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            float v = 0.f;
            if (r0 + i < m && p.t_len[lo + r0 + i] > 0) v = p.dz3[(int64_t)(r0 + i) * H1c + t];
            d3s[i][t] = v;
            s += v;
        }
        part[t] = s;  // db3 partial
    }
    __syncthreads();
    // dZ2 = (dZ3 W3) * dropout * (1 - a2^2): thread = (unit k of 128, half of the tile); W3 is [256 x 128] row-major
    {
        constexpr int RH = R / 2;
        const int k = t & (H2c - 1), g = t >> 7;
        float acc[RH];
#pragma unroll
        for (int i = 0; i < RH; ++i) acc[i] = 0.f;
        stream_matvec<RH, H1c, 32, H1c>(acc, p.W3 + k, H2c, d3s, g * RH);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            const int row = g * RH + i;
            float v = 0.f;
            if (r0 + row < m) {
                v = acc[i];
                if (p.drop.enabled) v *= dropout_factor(p.drop, r0 + row, k, H2c);
                const float av = p.a2[(int64_t)(r0 + row) * H2c + k];
                v *= 1.f - av * av;
                p.dz2[(int64_t)(r0 + row) * H2c + k] = v;
            }
            d2s[row][k] = v;
            s += v;
        }
        part[H1c + g * H2c + k] = s;  // db2 partial of this half
    }
    __syncthreads();
    // dZ1 = (dZ2 W2) * (1 - a1^2): thread = unit k of 256; W2 is [128 x 256] row-major
    {
        const int k = t;
        float acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = 0.f;
        stream_matvec<R, H2c, 32, H2c>(acc, p.W2 + k, H1c, d2s, 0);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (r0 + i < m) {
                const float av = p.a1[(int64_t)(r0 + i) * H1c + k];
                const float v = acc[i] * (1.f - av * av);
                p.dz1[(int64_t)(r0 + i) * H1c + k] = v;
                s += v;
            }
        }
        part[H1c + 2 * H2c + k] = s;  // db1 partial
    }
}

// ------------------------------------------------------------------------------------------------ 3b / 4c segments
// One warp per chunk of <= kSegChunk sorted entries of one (batch, column) segment; coefficient and source row of an
// entry are read in SORTED order (plan-time gather), the chunk's bounds from one int4.
__device__ __forceinline__ void seg_chunks_body(const FusedSeg& s, const float* __restrict__ src,
                                                float* __restrict__ grad, float* __restrict__ bias_grad, int warp,
                                                int n_warps) {
    constexpr int VEC = 2, W = H1c;
    const int lane = threadIdx.x & 31;
    const int c_lo = s.batch_chunk_off[s.b], c_hi = s.batch_chunk_off[s.b + 1];
    for (int c = c_lo + warp; c < c_hi; c += n_warps) {
        const int4 mt = s.meta[c];
        const int e0 = mt.x, e1 = mt.y, row_out = mt.z;
        const int k = mt.w & 0xffff, n_ch = mt.w >> 16;
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        float bsum = 0.f;
        for (int eb = e0; eb < e1; eb += 32) {
            const int e = eb + lane;
            float c_l = 0.f;
            int r_l = 0;
            if (e < e1) {
                c_l = s.coef_sorted[e];
                r_l = s.row_sorted[e];
            }
            bsum += c_l;
            const int cnt = min(32, e1 - eb);
            for (int i = 0; i < cnt; i += 4) {  // four source rows in flight; slots past cnt weigh 0
                float cc[4];
                float4 x[4][VEC];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int sl = min(i + q, cnt - 1);
                    const float cq = __shfl_sync(0xffffffffu, c_l, sl);
                    const int rq = __shfl_sync(0xffffffffu, r_l, sl);
                    cc[q] = (i + q < cnt) ? cq : 0.f;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[q][v] = ld4(src + (int64_t)rq * W + v * 128 + lane * 4);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) fma4(acc[v], cc[q], x[q][v]);
            }
        }
        bsum = warp_sum(bsum);
        if (n_ch == 1) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) st4(grad + (int64_t)row_out * W + v * 128 + lane * 4, acc[v]);
            if (bias_grad != nullptr && lane == 0) bias_grad[row_out] = bsum;
        } else {
            const int64_t slot = c - c_lo;
#pragma unroll
            for (int v = 0; v < VEC; ++v) st4(s.part + slot * W + v * 128 + lane * 4, acc[v]);
            if (lane == 0) s.part_bias[slot] = bsum;
            __threadfence();
            __syncwarp();
            const int64_t first = slot - k;
            int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&s.cnt[first], 1);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket == n_ch - 1) {  // last chunk of the segment: add the partial rows in chunk order
                __threadfence();
                float4 tot[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) tot[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                float bt = 0.f;
                for (int q = 0; q < n_ch; ++q) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) add4(tot[v], ldcg4(s.part + (first + q) * W + v * 128 + lane * 4));
                    if (lane == 0) bt += __ldcg(s.part_bias + first + q);
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) st4(grad + (int64_t)row_out * W + v * 128 + lane * 4, tot[v]);
                if (lane == 0) {
                    if (bias_grad != nullptr) bias_grad[row_out] = bt;
                    s.cnt[first] = 0;
                }
            }
        }
    }
}

// 3b' / 4c': the segmented reductions with the source rows streamed through a warp's bulk-copy ring (bulk.cuh)
__device__ __forceinline__ void seg_chunks_bulk_body(const FusedSeg& s, const float* __restrict__ src,
                                                     float* __restrict__ grad, float* __restrict__ bias_grad,
                                                     bulk::WarpRing<kBulkSlots>& rg, int warp, int n_warps) {
    constexpr int W = H1c;
    const int lane = threadIdx.x & 31;
    const int c_lo = s.batch_chunk_off[s.b], c_hi = s.batch_chunk_off[s.b + 1];
    for (int c = c_lo + warp; c < c_hi; c += n_warps) {
        const int4 mt = s.meta[c];
        const int e0 = mt.x, e1 = mt.y, row_out = mt.z;
        const int k = mt.w & 0xffff, n_ch = mt.w >> 16;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        float bsum = 0.f;
        for (int eb = e0; eb < e1; eb += 32) {
            const int e = eb + lane;
            float c_l = 0.f;
            int r_l = 0;
            if (e < e1) {
                c_l = s.coef_sorted[e];
                r_l = s.row_sorted[e];
            }
            bsum += c_l;
            bulk::gather_rows(rg, src, r_l, min(32, e1 - eb), lane, [&](int t0, int nv, const float4(&x)[4][2]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < nv) {
                        const float cq = __shfl_sync(0xffffffffu, c_l, t0 + q);
                        fma4(acc0, cq, x[q][0]);
                        fma4(acc1, cq, x[q][1]);
                    }
                }
            });
        }
        bsum = warp_sum(bsum);
        if (n_ch == 1) {
            st4(grad + (int64_t)row_out * W + lane * 4, acc0);
            st4(grad + (int64_t)row_out * W + 128 + lane * 4, acc1);
            if (bias_grad != nullptr && lane == 0) bias_grad[row_out] = bsum;
        } else {
            const int64_t slot = c - c_lo;
            st4(s.part + slot * W + lane * 4, acc0);
            st4(s.part + slot * W + 128 + lane * 4, acc1);
            if (lane == 0) s.part_bias[slot] = bsum;
            __threadfence();
            __syncwarp();
            const int64_t first = slot - k;
            int ticket = 0;
            if (lane == 0) ticket = atomicAdd(&s.cnt[first], 1);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket == n_ch - 1) {  // last chunk of the segment: add the partial rows in chunk order
                __threadfence();
                float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
                float bt = 0.f;
                for (int q = 0; q < n_ch; ++q) {
                    add4(t0, ldcg4(s.part + (first + q) * W + lane * 4));
                    add4(t1, ldcg4(s.part + (first + q) * W + 128 + lane * 4));
                    if (lane == 0) bt += __ldcg(s.part_bias + first + q);
                }
                st4(grad + (int64_t)row_out * W + lane * 4, t0);
                st4(grad + (int64_t)row_out * W + 128 + lane * 4, t1);
                if (lane == 0) {
                    if (bias_grad != nullptr) bias_grad[row_out] = bt;
                    s.cnt[first] = 0;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ 4a dW3 / dW2 tiles
// out[n][k] = sum_r Dn[r][n] * Xk[r][k] over the batch rows r (both operands row-major with r outermost: coalesced).
// 64 x 64 output tile per block, one slice of <= kDwSlice rows; 256 threads x (4 x 4) outputs.
__device__ __forceinline__ void dw_tile_body(const FusedGrad& p, int role) {
    __shared__ __align__(16) float As[kDwSlice][64];
    __shared__ __align__(16) float Bs[kDwSlice][64];
    __shared__ int s_last;
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int m = hi - lo;
    // role: [0, 64) dW3 (256 x 128 -> 4 x 2 tiles x 8 slices), [64, 128) dW2 (128 x 256 -> 2 x 4 tiles x 8 slices)
    const int which = role >> 6;
    const int idx = role & 63;
    const int tile = idx >> 3, slice = idx & 7;
    const int n_slices = (m + kDwSlice - 1) / kDwSlice;
    if (slice >= n_slices) return;
    const float* Dn = which == 0 ? p.dz3 : p.dz2;
    const float* Xk = which == 0 ? p.c : p.a1;
    const int N = which == 0 ? H1c : H2c, K = which == 0 ? H2c : H1c;  // output is [N x K]
    const int tiles_k = K / 64;
    const int n0 = (tile / tiles_k) * 64, k0 = (tile % tiles_k) * 64;
    const int rs = slice * kDwSlice, re = min(m, rs + kDwSlice);
    const int t = threadIdx.x;
    // rows without targets carry no dZ3 (never written by the decoder): zero them here like the row kernel does
    for (int i = t; i < (re - rs) * 16; i += 256) {
        const int r = i >> 4, q = (i & 15) * 4;
        float4 a = ld4(Dn + (int64_t)(rs + r) * N + n0 + q);
        if (which == 0 && p.t_len[lo + rs + r] == 0) a = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&As[r][q]) = a;
        *reinterpret_cast<float4*>(&Bs[r][q]) = ld4(Xk + (int64_t)(rs + r) * K + k0 + q);
    }
    __syncthreads();
    const int tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int nr = re - rs;
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    float* out = p.G + (which == 0 ? p.oW3 : p.oW2);
    if (n_slices == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            st4(out + (int64_t)(n0 + ty * 4 + i) * K + k0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        return;
    }
    const int tile_id = which * 8 + tile;
    float* part = p.part_dw + ((int64_t)tile_id * 8 + slice) * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        st4(part + (ty * 4 + i) * 64 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(&p.dw_cnt[tile_id], 1) == n_slices - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = p.part_dw + (int64_t)tile_id * 8 * 4096;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < n_slices; ++q) add4(tot, ldcg4(base + q * 4096 + (ty * 4 + i) * 64 + tx * 4));
        st4(out + (int64_t)(n0 + ty * 4 + i) * K + k0 + tx * 4, tot);
    }
    if (t == 0) p.dw_cnt[tile_id] = 0;
}

// 4b: bias gradients db3 / db2 / db1 = sum over the row-kernel CTAs of their column-sum partials, in CTA order
__device__ __forceinline__ void db_finish_body(const FusedGrad& p) {
    int lo, hi;
    if (!batch_range(p.br, lo, hi)) return;
    const int n_cta = (hi - lo + p.rows_per_cta - 1) / p.rows_per_cta;
    const int t = threadIdx.x;
    float s3 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int q = 0; q < n_cta; ++q) {
        const float* part = p.part_db + (int64_t)q * kDbPartStride;
        s3 += part[t];
        s1 += part[H1c + 2 * H2c + t];
        if (t < H2c) s2 += part[H1c + t] + part[H1c + H2c + t];
    }
    p.G[p.ob3 + t] = s3;
    p.G[p.ob1 + t] = s1;
    if (t < H2c) p.G[p.ob2 + t] = s2;
}

// ------------------------------------------------------------------------------------------------ kernels
template <int R>
__global__ void __launch_bounds__(256) ae_fwd_rows_kernel(FusedFwd p) {
    DMT_PDL_ENTRY();
    fwd_rows_body<R>(p, blockIdx.x);
}

__global__ void __launch_bounds__(256) ae_dec_chunks_kernel(FusedDec p) {
    DMT_PDL_ENTRY();
    dec_chunks_body(p);
}

__global__ void __launch_bounds__(128, 4) ae_dec_chunks4_kernel(FusedDec p) {
    DMT_PDL_ENTRY();
    dec_chunks4_body(p);
}

__global__ void __launch_bounds__(kBulkWarps * 32) ae_dec_chunks_bulk_kernel(FusedDec p) {
    DMT_PDL_ENTRY();
    dec_chunks_bulk_body(p);
}

__global__ void __launch_bounds__(kBulkWarps * 32) ae_seg_chunks_bulk_kernel(FusedSeg s, const float* src, float* grad,
                                                                             float* bias_grad) {
    __shared__ __align__(128) float ring[kBulkWarps][kBulkSlots][H1c];
    __shared__ __align__(8) uint64_t bars[kBulkWarps][kBulkSlots];
    if (s.active != nullptr && s.active[s.b] == 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bulk::WarpRing<kBulkSlots> rg = bulk::ring_setup<kBulkSlots>(&ring[wid][0][0], &bars[wid][0], lane);
    __syncthreads();
    seg_chunks_bulk_body(s, src, grad, bias_grad, rg, blockIdx.x * kBulkWarps + wid, gridDim.x * kBulkWarps);
}

// 3a and 3b are separate kernels on parallel branches of the step graph: the row kernel keeps 64 weights per thread in
// flight (~100 registers), the segment kernel needs 64 registers and four resident blocks per SM.
template <int R>
__global__ void __launch_bounds__(256) ae_bwd_rows_kernel(FusedBwd p) {
    DMT_PDL_ENTRY();
    bwd_rows_body<R>(p, blockIdx.x);
}

__global__ void __launch_bounds__(256) ae_seg_chunks_kernel(FusedSeg s, const float* src, float* grad,
                                                            float* bias_grad) {
    if (s.active != nullptr && s.active[s.b] == 0) return;
    seg_chunks_body(s, src, grad, bias_grad, blockIdx.x * 8 + (threadIdx.x >> 5), gridDim.x * 8);
}

__global__ void __launch_bounds__(256) ae_grad_phase_kernel(FusedGrad p, FusedSeg s, const float* src, float* grad) {
    DMT_PDL_ENTRY();
    const int blk = blockIdx.x;
    if (blk < 128) {
        dw_tile_body(p, blk);
        return;
    }
    if (blk == 128) {
        db_finish_body(p);
        return;
    }
    if (s.active != nullptr && s.active[s.b] == 0) return;
    seg_chunks_body(s, src, grad, nullptr, (blk - 129) * 8 + (threadIdx.x >> 5), (gridDim.x - 129) * 8);
}

// 5: sum of squares of the gradient, one partial per block (no tail: the Adam blocks add the partials themselves).
// Block 0 also advances the step counter (every reader of the old value — the dropout draws of this step — ran in
// earlier launches), block 1 reduces the batch loss.
__global__ void __launch_bounds__(256) norm_prepare_kernel(const float* __restrict__ g, int64_t n, float* partial,
                                                           AdamScalars* sc, int* step_dev, const float* loss_rows,
                                                           const int32_t* t_len, const int32_t* n_targets_ptr,
                                                           float* loss_out, BatchRef br) {
    __shared__ float sh[32];
    DMT_PDL_ENTRY();
    int lo, hi;
    if (!batch_range(br, lo, hi)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->active = 0;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *step_dev += 1;
        sc->active = 1;
    }
    float s = 0.f;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = ld4(g + 4 * i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += g[i] * g[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    if (blockIdx.x == 1 % gridDim.x && loss_out != nullptr) {
        float l = 0.f;
        for (int i = threadIdx.x; i < hi - lo; i += blockDim.x)
            if (t_len[lo + i] > 0) l += loss_rows[i];
        l = block_sum(l, sh);
        if (threadIdx.x == 0) loss_out[0] = l / (float)n_targets_ptr[0];
    }
}

// 6: dense Adam(+L2). Every block first adds the norm partials in a fixed order (same value in every block), derives
// the clip coefficient and the bias corrections from the step counter, then updates its slice; it also writes the
// transposed shadows of the W2 / W3 elements it updates.
__global__ void __launch_bounds__(256) adam_shadow_kernel(float* __restrict__ w, float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                          const AdamScalars* __restrict__ sc, AdamHyper hp,
                                                          const float* __restrict__ partial, int n_partial,
                                                          const int* __restrict__ step_dev, int64_t oW2, int64_t oW3,
                                                          float* __restrict__ W2t, float* __restrict__ W3t) {
    __shared__ float sh[32];
    __shared__ float s_sc[3];
    DMT_PDL_ENTRY();
    if (sc->active == 0) return;
    {
        float tot = 0.f;
        for (int i = threadIdx.x; i < n_partial; i += blockDim.x) tot += partial[i];
        tot = block_sum(tot, sh);
        if (threadIdx.x == 0) {
            const int64_t t = *step_dev;
            float coef = 1.f;
            if (hp.max_norm > 0.f) coef = fminf(1.f, hp.max_norm / (sqrtf(tot) + 1e-6f));
            const double bc1 = 1.0 - pow(hp.beta1, (double)t);
            const double bc2 = 1.0 - pow(hp.beta2, (double)t);
            s_sc[0] = coef;
            s_sc[1] = (float)(hp.lr / bc1);
            s_sc[2] = (float)sqrt(bc2);
        }
        __syncthreads();
    }
    const float coef = s_sc[0], step_size = s_sc[1], bc2_sqrt = s_sc[2];
    const float b2 = (float)hp.beta2, eps = (float)hp.eps, wd = (float)hp.weight_decay;
    const float omb1 = (float)(1.0 - hp.beta1), omb2 = (float)(1.0 - hp.beta2);
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 w4 = ld4(w + 4 * i), g4 = ld4(g + 4 * i), m4 = ld4(m + 4 * i), v4 = ld4(v + 4 * i);
        float* pw = &w4.x;
        float* pg = &g4.x;
        float* pm = &m4.x;
        float* pv = &v4.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gg = pg[k] * coef + wd * pw[k];
            pm[k] = pm[k] + (gg - pm[k]) * omb1;
            pv[k] = pv[k] * b2 + omb2 * gg * gg;
            const float denom = sqrtf(pv[k]) / bc2_sqrt + eps;
            pw[k] = pw[k] - step_size * (pm[k] / denom);
        }
        st4(w + 4 * i, w4);
        st4(m + 4 * i, m4);
        st4(v + 4 * i, v4);
        st4(g + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        const int64_t base = 4 * i;
        if (base >= oW2 && base < oW2 + (int64_t)H2c * H1c) {  // W2[n][k] (128 x 256) -> W2t[k][n]
            const int e = (int)(base - oW2), nn = e / H1c, kk = e % H1c;
#pragma unroll
            for (int q = 0; q < 4; ++q) W2t[(kk + q) * H2c + nn] = pw[q];
        } else if (base >= oW3 && base < oW3 + (int64_t)H1c * H2c) {  // W3[n][k] (256 x 128) -> W3t[k][n]
            const int e = (int)(base - oW3), nn = e / H2c, kk = e % H2c;
#pragma unroll
            for (int q = 0; q < 4; ++q) W3t[(kk + q) * H1c + nn] = pw[q];
        }
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gg = g[i] * coef + wd * w[i];
        const float mm = m[i] + (gg - m[i]) * omb1;
        const float vv = v[i] * b2 + omb2 * gg * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        w[i] = w[i] - step_size * (mm / denom);
        m[i] = mm;
        v[i] = vv;
        g[i] = 0.f;
    }
}

__global__ void __launch_bounds__(256) shadow_refresh_kernel(const float* __restrict__ W2, const float* __restrict__ W3,
                                                             float* __restrict__ W2t, float* __restrict__ W3t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 2 x 32768 elements
    if (i < H2c * H1c) {
        const int nn = i / H1c, kk = i % H1c;
        W2t[kk * H2c + nn] = W2[i];
    } else if (i < 2 * H2c * H1c) {
        const int e = i - H2c * H1c, nn = e / H2c, kk = e % H2c;
        W3t[kk * H1c + nn] = W3[e];
    }
}

// ------------------------------------------------------------------------------------------------ plan kernels
__global__ void plan_dec_meta_kernel(const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_row,
                                     int n_rows, const int32_t* __restrict__ rows, const int32_t* __restrict__ row_off,
                                     const int32_t* __restrict__ row_batch, const int32_t* __restrict__ t_indptr,
                                     const int32_t* __restrict__ ent_off, int4* __restrict__ meta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chunk_off[n_rows]) return;
    const int j = chunk_row[c];
    const int k = c - chunk_off[j];
    const int u = rows[j];
    const int r0 = t_indptr[u], len = t_indptr[u + 1] - r0;
    const int n_ch = (len + kDecChunk - 1) / kDecChunk;
    const int cnt = min(kDecChunk, len - k * kDecChunk);
    const int jl = j - row_off[row_batch[j]];
    meta[c] = make_int4(jl, r0 + k * kDecChunk, ent_off[j] + k * kDecChunk, cnt | (k << 8) | (n_ch << 20));
}

__global__ void plan_seg_meta_kernel(const int32_t* __restrict__ n_seg, const int32_t* __restrict__ seg_off,
                                     const int32_t* __restrict__ seg_key, const int32_t* __restrict__ seg_chunk_off,
                                     int n_cols, int4* __restrict__ meta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg[0]) return;
    const int c0 = seg_chunk_off[s], c1 = seg_chunk_off[s + 1];
    const int e0 = seg_off[s], e1 = seg_off[s + 1];
    const int col = (int)((uint32_t)seg_key[s] % (uint32_t)n_cols);
    for (int c = c0; c < c1; ++c) {
        const int k = c - c0;
        const int a = e0 + k * kSegChunk;
        meta[c] = make_int4(a, min(e1, a + kSegChunk), col, k | ((c1 - c0) << 16));
    }
}

// sorted-order copies: row_sorted[e] = ent_row[perm[e]], val_sorted[e] = val_ord[perm[e]], inv_perm[perm[e]] = e
__global__ void plan_sorted_kernel(const int32_t* __restrict__ perm, int64_t n, const int32_t* __restrict__ ent_row,
                                   const float* __restrict__ val_ord, int32_t* __restrict__ row_sorted,
                                   float* __restrict__ val_sorted, int32_t* __restrict__ inv_perm) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int id = perm[e];
        row_sorted[e] = ent_row[id];
        if (val_sorted != nullptr) val_sorted[e] = val_ord[id];
        if (inv_perm != nullptr) inv_perm[id] = (int32_t)e;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ launchers
int launch_fused_fwd(const FusedFwd& p, int n_rows_max, cudaStream_t st, bool pdl, int R) {
    if (n_rows_max <= 0) return 0;
    const int grid = (n_rows_max + R - 1) / R;
    if (R == 4) {
        DMT_CUDA(launch_k(ae_fwd_rows_kernel<4>, grid, 256, 0, st, pdl, p));
    } else if (R == 16) {
        DMT_CUDA(launch_k(ae_fwd_rows_kernel<16>, grid, 256, 0, st, pdl, p));
    } else {
        DMT_REQUIRE(R == kFusedRows, "launch_fused_fwd: row tile must be 4, 8 or 16");
        DMT_CUDA(launch_k(ae_fwd_rows_kernel<kFusedRows>, grid, 256, 0, st, pdl, p));
    }
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_fused_dec(const FusedDec& p, int blocks_hint, int gather, cudaStream_t st, bool pdl) {
    if (gather == 1) {  // bulk-copy rings: five 128-thread blocks per SM
        const int blocks = blocks_hint > 0 ? blocks_hint : kNumSMs * 5;
        DMT_CUDA(launch_k(ae_dec_chunks_bulk_kernel, blocks, kBulkWarps * 32, 0, st, pdl, p));
        DMT_LAUNCH_CHECK();
        return 0;
    }
    if (gather == 2) {  // four-warp blocks with next-chunk prefetch; blocks_hint counts 256-thread blocks
        const int blocks = blocks_hint > 0 ? blocks_hint * 2 : kNumSMs * 4;
        DMT_CUDA(launch_k(ae_dec_chunks4_kernel, blocks, 128, 0, st, pdl, p));
        DMT_LAUNCH_CHECK();
        return 0;
    }
    const int blocks = blocks_hint > 0 ? blocks_hint : kNumSMs * 2;
    DMT_CUDA(launch_k(ae_dec_chunks_kernel, blocks, 256, 0, st, pdl, p));
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_fused_bwd_rows(const FusedBwd& p, int n_rows_max, cudaStream_t st, bool pdl, int R) {
    if (n_rows_max <= 0) return 0;
    const int grid = (n_rows_max + R - 1) / R;
    if (R == 4) {
        DMT_CUDA(launch_k(ae_bwd_rows_kernel<4>, grid, 256, 0, st, pdl, p));
    } else if (R == 16) {
        DMT_CUDA(launch_k(ae_bwd_rows_kernel<16>, grid, 256, 0, st, pdl, p));
    } else {
        DMT_REQUIRE(R == kFusedRows, "launch_fused_bwd_rows: row tile must be 4, 8 or 16");
        DMT_CUDA(launch_k(ae_bwd_rows_kernel<kFusedRows>, grid, 256, 0, st, pdl, p));
    }
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_fused_seg_chunks(const FusedSeg& s, const float* src, float* grad, float* bias_grad, int n_chunk_max,
                            int gather, cudaStream_t st) {
    if (gather == 1) {
        int blocks = (n_chunk_max + kBulkWarps - 1) / kBulkWarps;
        if (blocks > kNumSMs * 6) blocks = kNumSMs * 6;
        if (blocks < 1) blocks = 1;
        ae_seg_chunks_bulk_kernel<<<blocks, kBulkWarps * 32, 0, st>>>(s, src, grad, bias_grad);
        DMT_LAUNCH_CHECK();
        return 0;
    }
    int seg_blocks = (n_chunk_max + 7) / 8;
    if (seg_blocks > kNumSMs * 2) seg_blocks = kNumSMs * 2;
    if (seg_blocks < 1) seg_blocks = 1;
    ae_seg_chunks_kernel<<<seg_blocks, 256, 0, st>>>(s, src, grad, bias_grad);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_fused_grad_phase(const FusedGrad& p, const FusedSeg& s, const float* src, float* grad, int n_chunk_max,
                            cudaStream_t st, bool pdl) {
    int seg_blocks = (n_chunk_max + 7) / 8;
    if (seg_blocks > kNumSMs) seg_blocks = kNumSMs;
    if (seg_blocks < 1) seg_blocks = 1;
    DMT_CUDA(launch_k(ae_grad_phase_kernel, 129 + seg_blocks, 256, 0, st, pdl, p, s, src, grad));
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_norm_prepare(const float* g, int64_t n, float* partial, AdamScalars* sc, int* step_dev,
                        const float* loss_rows, const int32_t* t_len, const int32_t* n_targets_ptr, float* loss_out,
                        BatchRef br, cudaStream_t st, bool pdl) {
    DMT_CUDA(launch_k(norm_prepare_kernel, kNormBlocks, 256, 0, st, pdl, g, n, partial, sc, step_dev, loss_rows, t_len,
                      n_targets_ptr, loss_out, br));
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_adam_shadow(float* w, float* g, float* m, float* v, int64_t n, const AdamScalars* sc, AdamHyper hp,
                       const float* partial, const int* step_dev, int64_t oW2, int64_t oW3, float* W2t, float* W3t,
                       cudaStream_t st, bool pdl) {
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    DMT_CUDA(launch_k(adam_shadow_kernel, (int)blocks, 256, 0, st, pdl, w, g, m, v, n, sc, hp, partial, (int)kNormBlocks,
                      step_dev, oW2, oW3, W2t, W3t));
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_shadow_refresh(const float* W2, const float* W3, float* W2t, float* W3t, cudaStream_t st) {
    shadow_refresh_kernel<<<(2 * H2c * H1c + 255) / 256, 256, 0, st>>>(W2, W3, W2t, W3t);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_plan_fused(const FusedPlanArgs& a, cudaStream_t st) {
    if (a.dec_chunk_cap > 0) {
        plan_dec_meta_kernel<<<(int)((a.dec_chunk_cap + 255) / 256), 256, 0, st>>>(
            a.t_chunk_off, a.t_chunk_row, a.n_rows, a.rows, a.row_off, a.row_batch, a.t_indptr, a.t_ent_off, a.dec_meta);
        DMT_LAUNCH_CHECK();
    }
    for (int side = 0; side < 2; ++side) {
        const FusedPlanSide& s = side == 0 ? a.t : a.d;
        if (s.n_entries <= 0) continue;
        plan_seg_meta_kernel<<<(int)((s.n_entries + 255) / 256), 256, 0, st>>>(s.n_seg, s.seg_off, s.seg_key,
                                                                              s.seg_chunk_off, s.n_cols, s.seg_meta);
        DMT_LAUNCH_CHECK();
        int blocks = (int)((s.n_entries + 1023) / 1024);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        plan_sorted_kernel<<<blocks, 256, 0, st>>>(s.perm, s.n_entries, s.ent_row, s.val_ord, s.row_sorted, s.val_sorted,
                                                   s.inv_perm);
        DMT_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace dmt
