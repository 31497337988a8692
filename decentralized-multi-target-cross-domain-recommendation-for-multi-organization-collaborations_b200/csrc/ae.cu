// AAE sparse ends (reference src/models/ae.py:98-157):
//   encoder first layer  = CSR SpMM over the batch rows with bias + tanh fused (ae.py:101-110)
//   decoder last layer   = row-gather SDDMM  o_f = A3[r_f] . W4[c_f] + b4[c_f]  fused with the loss, its
//                          derivative, and the first backward product dZ3 = (G W4) * (1 - A3^2)  (ae.py:135-156)
// Both read every weight row with 128-bit coalesced loads (one 1 KB row of H=256 floats = 2 float4 per lane).
// Roofline: HBM/L2 bytes — per target 4*H (W4 row) + 16 B; per data entry 4*H + 8 B.
#include "kernels.cuh"

namespace dmt {

// One block per batch row, one thread per hidden unit.
__device__ __forceinline__ void ae_encoder_fwd_body(const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices,
                                                            const float* __restrict__ val,
                                                            const float* __restrict__ W1t,
                                                            const float* __restrict__ b1, int H,
                                                            float* __restrict__ A1, BatchRef br) {
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    int j = blockIdx.x;
    if (j >= hi - lo) return;
    int u = rows[lo + j];
    int e0 = indptr[u], e1 = indptr[u + 1];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float acc = 0.f;
        int e = e0;
        for (; e + 4 <= e1; e += 4) {  // 4 independent row loads in flight
            int c0 = indices[e], c1 = indices[e + 1], c2 = indices[e + 2], c3 = indices[e + 3];
            float v0 = val[e], v1 = val[e + 1], v2 = val[e + 2], v3 = val[e + 3];
            float w0 = W1t[(int64_t)c0 * H + h], w1 = W1t[(int64_t)c1 * H + h];
            float w2 = W1t[(int64_t)c2 * H + h], w3 = W1t[(int64_t)c3 * H + h];
            acc = fmaf(v0, w0, acc);
            acc = fmaf(v1, w1, acc);
            acc = fmaf(v2, w2, acc);
            acc = fmaf(v3, w3, acc);
        }
        for (; e < e1; ++e) acc = fmaf(val[e], W1t[(int64_t)indices[e] * H + h], acc);
        A1[(int64_t)j * H + h] = tanhf(acc + b1[h]);
    }
}

__global__ void __launch_bounds__(512) ae_encoder_fwd_kernel(const int32_t* rows, const int32_t* indptr,
                                                            const int32_t* indices, const float* val,
                                                            const float* W1t, const float* b1, int H, float* A1,
                                                            BatchRef br) {
    ae_encoder_fwd_body(rows, indptr, indices, val, W1t, b1, H, A1, br);
}
__global__ void __launch_bounds__(512) ae_encoder_fwd_group(const OrgDev* __restrict__ orgs, int b, int H) {
    const OrgDev& o = orgs[blockIdx.z];
    ae_encoder_fwd_body(o.rows, o.d_indptr, o.d_indices, o.d_val, o.P + o.oW1, o.P + o.ob1, H, o.a1,
                        BatchRef{o.row_off, o.active, b, 0, 0});
}

// One block (8 warps) per batch row. VEC = H / 128: each lane owns VEC float4 slices of the hidden vector.
template <int VEC>
__global__ void __launch_bounds__(256) ae_decoder_fwd_kernel(const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices,
                                                            const float* __restrict__ target,
                                                            const float* __restrict__ A3,
                                                            const float* __restrict__ W4,
                                                            const float* __restrict__ b4, int loss_kind,
                                                            const int32_t* __restrict__ n_targets,
                                                            const int32_t* __restrict__ ent_off,
                                                            float* __restrict__ pred, float* __restrict__ gout,
                                                            float* __restrict__ dZ3, float* __restrict__ loss_rows,
                                                            int tanh_deriv, BatchRef br) {
    constexpr int H = VEC * 128;
    __shared__ float s_acc[8][H];
    __shared__ float s_loss[8];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    int j = blockIdx.x;
    if (j >= hi - lo) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int u = rows[lo + j];
    const int e0 = indptr[u], e1 = indptr[u + 1];
    const bool train = gout != nullptr;
    const float inv_n = train ? 1.f / (float)n_targets[br.row_off ? br.b : 0] : 0.f;
    const int64_t out_base = ent_off ? (int64_t)ent_off[lo + j] - e0 : 0;  // entry e is written at out_base + e
    float4 a[VEC], acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        a[v] = ld4(A3 + (int64_t)j * H + v * 128 + lane * 4);
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float loss_acc = 0.f;
    // each warp takes 32-entry chunks; the chunk's columns/targets are loaded coalesced, then broadcast
    for (int eb = e0 + wid * 32; eb < e1; eb += 8 * 32) {
        int e = eb + lane;
        int c_l = 0;
        float y_l = 0.f;
        if (e < e1) {
            c_l = indices[e];
            y_l = target ? target[e] : 0.f;
        }
        int cnt = min(32, e1 - eb);
        float o_l = 0.f;  // lane i keeps the prediction of entry eb+i
        int i = 0;
        for (; i + 2 <= cnt; i += 2) {  // two rows in flight per warp
            int c0 = __shfl_sync(0xffffffffu, c_l, i), c1 = __shfl_sync(0xffffffffu, c_l, i + 1);
            float4 w0[VEC], w1[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                w0[v] = ld4(W4 + (int64_t)c0 * H + v * 128 + lane * 4);
                w1[v] = ld4(W4 + (int64_t)c1 * H + v * 128 + lane * 4);
            }
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                d0 += a[v].x * w0[v].x + a[v].y * w0[v].y + a[v].z * w0[v].z + a[v].w * w0[v].w;
                d1 += a[v].x * w1[v].x + a[v].y * w1[v].y + a[v].z * w1[v].z + a[v].w * w1[v].w;
            }
            d0 = warp_sum(d0) + b4[c0];
            d1 = warp_sum(d1) + b4[c1];
            if (lane == i) o_l = d0;
            if (lane == i + 1) o_l = d1;
            if (train) {
                float g0 = loss_grad(loss_kind, d0, __shfl_sync(0xffffffffu, y_l, i)) * inv_n;
                float g1 = loss_grad(loss_kind, d1, __shfl_sync(0xffffffffu, y_l, i + 1)) * inv_n;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    acc[v].x += g0 * w0[v].x + g1 * w1[v].x;
                    acc[v].y += g0 * w0[v].y + g1 * w1[v].y;
                    acc[v].z += g0 * w0[v].z + g1 * w1[v].z;
                    acc[v].w += g0 * w0[v].w + g1 * w1[v].w;
                }
            }
        }
        for (; i < cnt; ++i) {
            int c0 = __shfl_sync(0xffffffffu, c_l, i);
            float4 w0[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) w0[v] = ld4(W4 + (int64_t)c0 * H + v * 128 + lane * 4);
            float d0 = 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) d0 += a[v].x * w0[v].x + a[v].y * w0[v].y + a[v].z * w0[v].z + a[v].w * w0[v].w;
            d0 = warp_sum(d0) + b4[c0];
            if (lane == i) o_l = d0;
            if (train) {
                float g0 = loss_grad(loss_kind, d0, __shfl_sync(0xffffffffu, y_l, i)) * inv_n;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    acc[v].x += g0 * w0[v].x;
                    acc[v].y += g0 * w0[v].y;
                    acc[v].z += g0 * w0[v].z;
                    acc[v].w += g0 * w0[v].w;
                }
            }
        }
        if (e < e1) {  // coalesced write-back of the chunk
            if (pred) pred[out_base + e] = o_l;
            if (train) {
                gout[out_base + e] = loss_grad(loss_kind, o_l, y_l) * inv_n;
                loss_acc += loss_value(loss_kind, o_l, y_l);
            }
        }
    }
    if (!train) return;
#pragma unroll
    for (int v = 0; v < VEC; ++v) st4(&s_acc[wid][v * 128 + lane * 4], acc[v]);
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) s_loss[wid] = loss_acc;
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_acc[w][h];
        float av = A3[(int64_t)j * H + h];
        dZ3[(int64_t)j * H + h] = tanh_deriv ? s * (1.f - av * av) : s;
    }
    if (threadIdx.x == 0) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) l += s_loss[w];
        loss_rows[j] = l;
    }
}

int launch_ae_encoder_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* val,
                          const float* W1t, const float* b1, int H, float* A1, int n_rows_max, BatchRef br,
                          cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    int threads = H >= 512 ? 512 : (H + 31) / 32 * 32;
    ae_encoder_fwd_kernel<<<n_rows_max, threads, 0, st>>>(rows, indptr, indices, val, W1t, b1, H, A1, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_ae_decoder_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                          const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                          const int32_t* n_targets, const int32_t* ent_off, float* pred, float* gout, float* dZ3,
                          float* loss_rows, int tanh_deriv, int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
#define DMT_DEC(V)                                                                                              \
    ae_decoder_fwd_kernel<V><<<n_rows_max, 256, 0, st>>>(rows, indptr, indices, target, A3, W4, b4, loss_kind, \
                                                         n_targets, ent_off, pred, gout, dZ3, loss_rows, tanh_deriv, br)
    if (H == 128) DMT_DEC(1);
    else if (H == 256) DMT_DEC(2);
    else if (H == 384) DMT_DEC(3);
    else if (H == 512) DMT_DEC(4);
    else {
        set_error("decoder hidden size must be 128, 256, 384 or 512");
        return DMT_E_ARG;
    }
#undef DMT_DEC
    DMT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- load-balanced training variant (engine)
// One block per CHUNK of <= kDecChunk targets of one batch row, so a heavy row (a user with ~2000 ratings) is spread
// over ~16 blocks instead of serialising one. Each block leaves a partial dZ3 row; the finish kernel adds a row's
// partials in chunk order (deterministic) and applies the tanh derivative.
template <int VEC>
__device__ __forceinline__ void ae_decoder_chunk_body(const int32_t* __restrict__ rows,
                                                              const int32_t* __restrict__ indptr,
                                                              const int32_t* __restrict__ indices,
                                                              const float* __restrict__ target,
                                                              const float* __restrict__ A3,
                                                              const float* __restrict__ W4,
                                                              const float* __restrict__ b4, int loss_kind,
                                                              const int32_t* __restrict__ n_targets,
                                                              const int32_t* __restrict__ ent_off, DecChunks dc,
                                                              float* __restrict__ gout, BatchRef br) {
    constexpr int H = VEC * 128;
    constexpr int PER_WARP = kDecChunk / 8;  // 16 targets per warp
    constexpr int GROUPS = PER_WARP / 4;     // in groups of four 1 KB weight rows
    // double-buffered so that a chunk needs ONE block barrier: the next chunk's partials go to the other buffer, and
    // nobody can be two chunks ahead of a thread that still reads (it would have to pass the barrier in between)
    __shared__ float s_acc[2][8][H];
    __shared__ float s_loss[2][8];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c_lo = dc.batch_chunk_off[br.b], c_hi = dc.batch_chunk_off[br.b + 1];
    const float inv_n = 1.f / (float)n_targets[br.b];
    int buf = 0;
    for (int c = c_lo + blockIdx.x; c < c_hi; c += gridDim.x, buf ^= 1) {
        const int j = dc.chunk_row[c];  // epoch-wide batch-row index
        const int u = rows[j];
        const int k = c - dc.chunk_off[j];
        const int r0 = indptr[u];
        const int e0 = r0 + k * kDecChunk;
        const int e1 = min(indptr[u + 1], e0 + kDecChunk);
        const int64_t out_base = (int64_t)ent_off[j] - r0;
        const float* a_row = A3 + (int64_t)(j - lo) * H;
        float4 a[VEC], acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            a[v] = ld4(a_row + v * 128 + lane * 4);
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float loss_acc = 0.f;
        const int eb = e0 + wid * PER_WARP;
        const int cnt = max(0, min(PER_WARP, e1 - eb));
        int c_l = 0;
        float y_l = 0.f;
        if (lane < cnt) {
            c_l = indices[eb + lane];
            y_l = target[eb + lane];
        }
        float o_l = 0.f;
        if (cnt > 0) {
            // software pipeline: the four weight rows of group g+1 are in flight while group g is reduced, so eight 1 KB
            // rows per warp are outstanding instead of a load-wait-compute sequence per group. Slots past `cnt` re-read
            // the last valid row and contribute with g = 0.
            float4 w[2][4][VEC];
            float bb[2][4];
            auto load_group = [&](int g, int slot) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = __shfl_sync(0xffffffffu, c_l, min(4 * g + q, cnt - 1));
#pragma unroll
                    for (int v = 0; v < VEC; ++v) w[slot][q][v] = ld4(W4 + (int64_t)col * H + v * 128 + lane * 4);
                    bb[slot][q] = b4[col];
                }
            };
            load_group(0, 0);
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                if (4 * g < cnt) {  // warp-uniform
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
                        const int t = 4 * g + q;
                        const bool valid = t < cnt;
                        d[q] = warp_sum(d[q]) + bb[slot][q];
                        if (lane == t && valid) o_l = d[q];
                        const float y = __shfl_sync(0xffffffffu, y_l, min(t, cnt - 1));
                        const float gq = valid ? loss_grad(loss_kind, d[q], y) * inv_n : 0.f;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            acc[v].x = fmaf(gq, w[slot][q][v].x, acc[v].x);
                            acc[v].y = fmaf(gq, w[slot][q][v].y, acc[v].y);
                            acc[v].z = fmaf(gq, w[slot][q][v].z, acc[v].z);
                            acc[v].w = fmaf(gq, w[slot][q][v].w, acc[v].w);
                        }
                    }
                }
            }
        }
        if (lane < cnt) {
            gout[out_base + eb + lane] = loss_grad(loss_kind, o_l, y_l) * inv_n;
            loss_acc = loss_value(loss_kind, o_l, y_l);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) st4(&s_acc[buf][wid][v * 128 + lane * 4], acc[v]);
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) s_loss[buf][wid] = loss_acc;
        __syncthreads();
        const int64_t slot_out = c - c_lo;
        for (int h = threadIdx.x; h < H; h += 256) {
            float s = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) s += s_acc[buf][w8][h];
            dc.dz_part[slot_out * H + h] = s;
        }
        if (threadIdx.x == 0) {
            float l = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) l += s_loss[buf][w8];
            dc.loss_part[slot_out] = l;
        }
    }
}

__device__ __forceinline__ void ae_decoder_finish_body(const float* __restrict__ A3, int H, DecChunks dc,
                                                                float* __restrict__ dZ3, float* __restrict__ loss_rows,
                                                                BatchRef br) {
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int jl = blockIdx.x;
    if (jl >= hi - lo) return;
    const int c_base = dc.batch_chunk_off[br.b];
    const int c0 = dc.chunk_off[lo + jl] - c_base, c1 = dc.chunk_off[lo + jl + 1] - c_base;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float s = 0.f;
        for (int c = c0; c < c1; ++c) s += dc.dz_part[(int64_t)c * H + h];
        const float av = A3[(int64_t)jl * H + h];
        dZ3[(int64_t)jl * H + h] = s * (1.f - av * av);
    }
    if (threadIdx.x == 0) {
        float l = 0.f;
        for (int c = c0; c < c1; ++c) l += dc.loss_part[c];
        loss_rows[jl] = l;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) ae_decoder_chunk_kernel(const int32_t* rows, const int32_t* indptr,
                                                              const int32_t* indices, const float* target,
                                                              const float* A3, const float* W4, const float* b4,
                                                              int loss_kind, const int32_t* n_targets,
                                                              const int32_t* ent_off, DecChunks dc, float* gout,
                                                              BatchRef br) {
    ae_decoder_chunk_body<VEC>(rows, indptr, indices, target, A3, W4, b4, loss_kind, n_targets, ent_off, dc, gout, br);
}
__global__ void __launch_bounds__(256) ae_decoder_finish_kernel(const float* A3, int H, DecChunks dc, float* dZ3,
                                                                float* loss_rows, BatchRef br) {
    ae_decoder_finish_body(A3, H, dc, dZ3, loss_rows, br);
}
template <int VEC>
__global__ void __launch_bounds__(256) ae_decoder_chunk_group(const OrgDev* __restrict__ orgs, int b) {
    const OrgDev& o = orgs[blockIdx.z];
    ae_decoder_chunk_body<VEC>(o.rows, o.t_indptr, o.t_indices, o.t_val, o.a3, o.P + o.oW4, o.P + o.ob4, DMT_LOSS_MSE,
                               o.t_batch_cnt, o.t_ent_off, o.dc, o.gbuf, BatchRef{o.row_off, o.active, b, 0, 0});
}
__global__ void __launch_bounds__(256) ae_decoder_finish_group(const OrgDev* __restrict__ orgs, int b, int H) {
    const OrgDev& o = orgs[blockIdx.z];
    ae_decoder_finish_body(o.a3, H, o.dc, o.dz3, o.loss_rows, BatchRef{o.row_off, o.active, b, 0, 0});
}

int launch_group_encoder(const OrgDev* orgs, int G, int b, int B, int H1, cudaStream_t st) {
    int threads = H1 >= 512 ? 512 : (H1 + 31) / 32 * 32;
    ae_encoder_fwd_group<<<dim3(B, 1, G), threads, 0, st>>>(orgs, b, H1);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_group_decoder(const OrgDev* orgs, int G, int b, int B, int H1, cudaStream_t st) {
    // chunks of all organizations share the persistent grid: x = chunk slots per organization, z = organization
    int per_org = (kNumSMs * 8 + G - 1) / G;
    if (per_org < 8) per_org = 8;
    dim3 grid(per_org, 1, G);
    if (H1 == 128) ae_decoder_chunk_group<1><<<grid, 256, 0, st>>>(orgs, b);
    else if (H1 == 256) ae_decoder_chunk_group<2><<<grid, 256, 0, st>>>(orgs, b);
    else if (H1 == 384) ae_decoder_chunk_group<3><<<grid, 256, 0, st>>>(orgs, b);
    else if (H1 == 512) ae_decoder_chunk_group<4><<<grid, 256, 0, st>>>(orgs, b);
    else { set_error("decoder hidden size must be 128, 256, 384 or 512"); return DMT_E_ARG; }
    DMT_LAUNCH_CHECK();
    ae_decoder_finish_group<<<dim3(B, 1, G), 256, 0, st>>>(orgs, b, H1);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_ae_decoder_chunks(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                             const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                             const int32_t* n_targets, const int32_t* ent_off, DecChunks dc, float* gout, float* dZ3,
                             float* loss_rows, int n_rows_max, BatchRef br, cudaStream_t st, int blocks_hint) {
    if (n_rows_max <= 0) return 0;
    // persistent over the batch's chunks (count known only on the device); 2 blocks per SM leaves room for the other
    // organizations' graphs that run concurrently on their own streams
    // default: two blocks per SM (the kernel's 128 registers x 256 threads fill the register file with two) — the
    // fastest for ONE organization; with many organizations per GPU fewer blocks leave room for the other
    // organizations' kernels on the same SMs and the round gets shorter although this kernel gets longer (blocks_hint)
    const int blocks = blocks_hint > 0 ? blocks_hint : kNumSMs * 2;
#define DMT_DECC(V)                                                                                             \
    ae_decoder_chunk_kernel<V><<<blocks, 256, 0, st>>>(rows, indptr, indices, target, A3, W4, b4, loss_kind,   \
                                                       n_targets, ent_off, dc, gout, br)
    if (H == 128) DMT_DECC(1);
    else if (H == 256) DMT_DECC(2);
    else if (H == 384) DMT_DECC(3);
    else if (H == 512) DMT_DECC(4);
    else {
        set_error("decoder hidden size must be 128, 256, 384 or 512");
        return DMT_E_ARG;
    }
#undef DMT_DECC
    DMT_LAUNCH_CHECK();
    ae_decoder_finish_kernel<<<n_rows_max, 256, 0, st>>>(A3, H, dc, dZ3, loss_rows, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int dmt_ae_encoder_fwd(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                       const float* val, const float* W1t, const float* b1, int H, float* A1, void* stream) {
    DMT_REQUIRE(n_rows >= 0 && H > 0, "dmt_ae_encoder_fwd: bad argument");
    return launch_ae_encoder_fwd(rows, indptr, indices, val, W1t, b1, H, A1, n_rows, batch_by_value(0, n_rows),
                                 as_stream(stream));
}

int dmt_ae_decoder_fwd(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                       const float* target, const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                       const int32_t* n_targets, float* pred, float* gout, float* dZ3, float* loss_rows,
                       int tanh_deriv, void* stream) {
    DMT_REQUIRE(n_rows >= 0, "dmt_ae_decoder_fwd: bad argument");
    DMT_REQUIRE(gout == nullptr || (n_targets && dZ3 && loss_rows && target), "dmt_ae_decoder_fwd: train mode needs "
                "n_targets, target, dZ3 and loss_rows");
    return launch_ae_decoder_fwd(rows, indptr, indices, target, A3, W4, b4, H, loss_kind, n_targets, nullptr, pred,
                                 gout, dZ3, loss_rows, tanh_deriv, n_rows, batch_by_value(0, n_rows),
                                 as_stream(stream));
}

}  // extern "C"
