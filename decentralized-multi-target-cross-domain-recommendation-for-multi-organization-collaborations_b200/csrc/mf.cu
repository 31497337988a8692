// MF / GMF rating kernels (reference src/models/mf.py:36-48,57-93; the GMF branch of src/models/nmf.py:122-126).
//   forward : one warp per rating; both embedding rows are read with 128-bit coalesced loads (H=128 -> one float4
//             per lane), the per-row bias is broadcast-added BEFORE the product exactly as the reference does,
//             then dot + global bias + loss + dloss/dpred in the same pass.
//   backward: sort-by-index (segments.cu) + one warp per unique row: g * (other row [+ side projection]) summed in
//             registers, written once -> the dense .grad of the embedding table without atomics.
// Roofline: HBM/L2 bytes, 2*(4H+4)+16 per rating forward.
#include "kernels.cuh"

namespace dmt {

constexpr int kMfBlocks = kNumSMs * 4;

template <int VEC>
__global__ void __launch_bounds__(256) mf_fwd_kernel(const int32_t* __restrict__ user, const int32_t* __restrict__ item,
                                                     const float* __restrict__ rating, int64_t n,
                                                     const float* __restrict__ Wu, const float* __restrict__ Wi,
                                                     const float* __restrict__ bu, const float* __restrict__ bi,
                                                     const float* __restrict__ bias, const float* __restrict__ pu,
                                                     const float* __restrict__ pi,
                                                     const float* __restrict__ colscale,
                                                     const float* __restrict__ add, int loss_kind,
                                                     float* __restrict__ pred, float* __restrict__ dpred,
                                                     float* __restrict__ q_out, float* __restrict__ partial) {
    constexpr int H = VEC * 128;
    __shared__ float s_l[8], s_g[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * 8 + wid, n_warps = (int64_t)gridDim.x * 8;
    const float b0 = bias ? bias[0] : 0.f;
    float loss_acc = 0.f, g_acc = 0.f;
    for (int64_t e = warp; e < n; e += n_warps) {
        const int u = user[e], i = item[e];
        const float bu_ = bu[u], bi_ = bi[i];
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int off = v * 128 + lane * 4;
            float4 a = ld4(Wu + (int64_t)u * H + off), b = ld4(Wi + (int64_t)i * H + off);
            a.x += bu_; a.y += bu_; a.z += bu_; a.w += bu_;
            b.x += bi_; b.y += bi_; b.z += bi_; b.w += bi_;
            float4 t = b;  // d(sum)/d(u~) = i~ [+ pu]; accumulated as u~ . (i~ + pu) + i~ . pi
            if (pu != nullptr) {
                float4 p = ld4(pu + e * H + off);
                t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
            }
            float4 q = make_float4(a.x * t.x, a.y * t.y, a.z * t.z, a.w * t.w);
            if (pi != nullptr) {
                float4 p = ld4(pi + e * H + off);
                q.x += b.x * p.x; q.y += b.y * p.y; q.z += b.z * p.z; q.w += b.w * p.w;
            }
            if (q_out != nullptr) st4(q_out + e * H + off, q);  // GMF vector (NCF: the input of the affine layer)
            if (colscale != nullptr) {
                float4 c = ld4(colscale + off);
                s += q.x * c.x + q.y * c.y + q.z * c.z + q.w * c.w;
            } else {
                s += q.x + q.y + q.z + q.w;
            }
        }
        s = warp_sum(s) + b0;
        if (add != nullptr) s += add[e];
        if (lane == 0) {
            pred[e] = s;
            const float y = rating[e];
            loss_acc += loss_value(loss_kind, s, y);
            const float g = loss_grad(loss_kind, s, y);
            g_acc += g;
            if (dpred != nullptr) dpred[e] = g;
        }
    }
    if (lane == 0) {
        s_l[wid] = loss_acc;
        s_g[wid] = g_acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f, g = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            l += s_l[w];
            g += s_g[w];
        }
        partial[blockIdx.x] = l;
        partial[gridDim.x + blockIdx.x] = g;
    }
}

__global__ void mf_finish_kernel(const float* __restrict__ partial, int nb, float* __restrict__ sums) {
    __shared__ float sh[32];
    float l = 0.f, g = 0.f;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        l += partial[i];
        g += partial[nb + i];
    }
    l = block_sum(l, sh);
    g = block_sum(g, sh);
    if (threadIdx.x == 0) {
        sums[0] = l;  // sum of per-rating losses
        sums[1] = g;  // sum of dloss/dpred (gradient of the global bias before the 1/n scale)
    }
}

// One warp per unique row (segment) of the table being differentiated.
template <int VEC>
__global__ void __launch_bounds__(256) mf_bwd_table_kernel(const int32_t* __restrict__ other, const float* __restrict__ Wo,
                                                           const float* __restrict__ bo,
                                                           const float* __restrict__ p_side,
                                                           const float* __restrict__ dpred, float scale,
                                                           const int32_t* __restrict__ perm,
                                                           const int32_t* __restrict__ seg_key,
                                                           const int32_t* __restrict__ seg_off,
                                                           const int32_t* __restrict__ n_seg,
                                                           const float* __restrict__ colscale,
                                                           float* __restrict__ dW, float* __restrict__ db) {
    constexpr int H = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
    const int ns = n_seg[0];
    for (int64_t s = warp; s < ns; s += n_warps) {
        const int e0 = seg_off[s], e1 = seg_off[s + 1];
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int eb = e0; eb < e1; eb += 32) {
            int e = eb + lane;
            int id_l = 0, o_l = 0;
            float g_l = 0.f;
            if (e < e1) {
                id_l = perm[e];
                o_l = other[id_l];
                g_l = dpred[id_l] * scale;
            }
            int cnt = min(32, e1 - eb);
            for (int i = 0; i < cnt; ++i) {
                const int id = __shfl_sync(0xffffffffu, id_l, i);
                const int o = __shfl_sync(0xffffffffu, o_l, i);
                const float g = __shfl_sync(0xffffffffu, g_l, i);
                const float b_ = bo[o];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int off = v * 128 + lane * 4;
                    float4 x = ld4(Wo + (int64_t)o * H + off);
                    x.x += b_; x.y += b_; x.z += b_; x.w += b_;
                    if (p_side != nullptr) {
                        float4 p = ld4(p_side + (int64_t)id * H + off);
                        x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w;
                    }
                    acc[v].x = fmaf(g, x.x, acc[v].x);
                    acc[v].y = fmaf(g, x.y, acc[v].y);
                    acc[v].z = fmaf(g, x.z, acc[v].z);
                    acc[v].w = fmaf(g, x.w, acc[v].w);
                }
            }
        }
        const int row = seg_key[s];
        float bsum = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (colscale != nullptr) {  // NCF: d/d(row) of sum_d a_d * row_d * other_d
                float4 c = ld4(colscale + v * 128 + lane * 4);
                acc[v].x *= c.x; acc[v].y *= c.y; acc[v].z *= c.z; acc[v].w *= c.w;
            }
            st4(dW + (int64_t)row * H + v * 128 + lane * 4, acc[v]);
            bsum += acc[v].x + acc[v].y + acc[v].z + acc[v].w;
        }
        // the bias is added to every hidden dim, so its gradient is the row-sum of the weight gradient
        bsum = warp_sum(bsum);
        if (lane == 0 && db != nullptr) db[row] = bsum;
    }
}

// d_p[e][:] = g_e * (W[idx[e]] + b[idx[e]])   (gradient w.r.t. a side-information projection, mf.py:82-90)
template <int VEC>
__global__ void __launch_bounds__(256) mf_bwd_side_kernel(const int32_t* __restrict__ idx, int64_t n,
                                                          const float* __restrict__ W, const float* __restrict__ b,
                                                          const float* __restrict__ dpred, float scale,
                                                          const float* __restrict__ colscale,
                                                          float* __restrict__ d_p) {
    constexpr int H = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
    for (int64_t e = warp; e < n; e += n_warps) {
        const int r = idx[e];
        const float g = dpred[e] * scale, b_ = b[r];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int off = v * 128 + lane * 4;
            float4 x = ld4(W + (int64_t)r * H + off);
            x.x = g * (x.x + b_); x.y = g * (x.y + b_); x.z = g * (x.z + b_); x.w = g * (x.w + b_);
            if (colscale != nullptr) {
                float4 c = ld4(colscale + off);
                x.x *= c.x; x.y *= c.y; x.z *= c.z; x.w *= c.w;
            }
            st4(d_p + e * H + off, x);
        }
    }
}

// out[e][col_off + d] = W[idx[e]][d] + b[idx[e]]  (embedding with the bias broadcast-added, written into one slice
// of the concatenated tower input; reference src/models/mlp.py:52-64,96; nmf.py:62-88,127)
template <int VEC>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int32_t* __restrict__ idx, int64_t n,
                                                        const float* __restrict__ W, const float* __restrict__ b,
                                                        float* __restrict__ out, int ld, int col_off) {
    constexpr int H = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
    for (int64_t e = warp; e < n; e += n_warps) {
        const int r = idx[e];
        const float b_ = b[r];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int off = v * 128 + lane * 4;
            float4 x = ld4(W + (int64_t)r * H + off);
            x.x += b_; x.y += b_; x.z += b_; x.w += b_;
            st4(out + e * ld + col_off + off, x);
        }
    }
}

// dW[r] = sum_{e in segment r} dOut[perm[e]][col_off : col_off + H],  db[r] = row-sum of dW[r]
template <int VEC>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ dOut, int ld, int col_off,
                                                        const int32_t* __restrict__ perm,
                                                        const int32_t* __restrict__ seg_key,
                                                        const int32_t* __restrict__ seg_off,
                                                        const int32_t* __restrict__ n_seg, float* __restrict__ dW,
                                                        float* __restrict__ db) {
    constexpr int H = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
    const int ns = n_seg[0];
    for (int64_t s = warp; s < ns; s += n_warps) {
        const int e0 = seg_off[s], e1 = seg_off[s + 1];
        float4 acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int eb = e0; eb < e1; eb += 32) {
            int id_l = (eb + lane < e1) ? perm[eb + lane] : 0;
            const int cnt = min(32, e1 - eb);
            for (int i = 0; i < cnt; ++i) {
                const int64_t id = __shfl_sync(0xffffffffu, id_l, i);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float4 x = ld4(dOut + id * ld + col_off + v * 128 + lane * 4);
                    acc[v].x += x.x; acc[v].y += x.y; acc[v].z += x.z; acc[v].w += x.w;
                }
            }
        }
        const int row = seg_key[s];
        float bsum = 0.f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            st4(dW + (int64_t)row * H + v * 128 + lane * 4, acc[v]);
            bsum += acc[v].x + acc[v].y + acc[v].z + acc[v].w;
        }
        bsum = warp_sum(bsum);
        if (lane == 0 && db != nullptr) db[row] = bsum;
    }
}

// out[d] = sum_e g[e] * scale * Q[e][d]  for d < width (any width): gradient of a 1-row weight (NCF affine) and of
// the GMF column scale. Stage 1: per-block partial over a slab of rows; stage 2: fixed-order sum of the partials.
constexpr int kWcsBlocks = kNumSMs * 2;
__global__ void __launch_bounds__(256) weighted_colsum_stage1(const float* __restrict__ g, float scale,
                                                              const float* __restrict__ Q, int64_t n, int width,
                                                              int ld, float* __restrict__ partial) {
    // thread t owns columns t, t+256, ...; rows strided over blocks
    for (int d = threadIdx.x; d < width; d += 256) {
        float s = 0.f;
        for (int64_t e = blockIdx.x; e < n; e += gridDim.x) s = fmaf(g[e] * scale, Q[e * ld + d], s);
        partial[(int64_t)blockIdx.x * width + d] = s;
    }
}
__global__ void weighted_colsum_stage2(const float* __restrict__ partial, int nb, int width, float* __restrict__ out) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= width) return;
    float s = 0.f;
    for (int i = 0; i < nb; ++i) s += partial[(int64_t)i * width + d];
    out[d] = s;
}

// loss + dloss/dpred of a finished prediction vector (MLP: the affine output), two-stage sums like mf_fwd
__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ y,
                                                       int64_t n, int kind, float* __restrict__ dpred,
                                                       float* __restrict__ partial) {
    __shared__ float sh[32];
    float l = 0.f, g = 0.f;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float o = pred[i], t = y[i];
        l += loss_value(kind, o, t);
        float d = loss_grad(kind, o, t);
        g += d;
        if (dpred != nullptr) dpred[i] = d;
    }
    l = block_sum(l, sh);
    g = block_sum(g, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = l;
        partial[gridDim.x + blockIdx.x] = g;
    }
}

}  // namespace dmt

using namespace dmt;

#define DMT_VEC_DISPATCH(H, CALL)                                     \
    do {                                                              \
        if ((H) == 128) { CALL(1); }                                  \
        else if ((H) == 256) { CALL(2); }                             \
        else if ((H) == 384) { CALL(3); }                             \
        else if ((H) == 512) { CALL(4); }                             \
        else { set_error("hidden size must be 128, 256, 384 or 512"); return DMT_E_ARG; } \
    } while (0)

extern "C" {

int64_t dmt_mf_scratch_floats(void) { return 2 * kMfBlocks; }

int dmt_mf_fwd(const int32_t* user, const int32_t* item, const float* rating, int64_t n, const float* Wu,
               const float* Wi, const float* bu, const float* bi, const float* bias, const float* pu, const float* pi,
               const float* colscale, const float* add, int H, int loss_kind, float* pred, float* dpred, float* q_out,
               float* sums, float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0 && pred && sums && scratch, "dmt_mf_fwd: bad argument");
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n + 7) / 8;
    if (blocks > kMfBlocks) blocks = kMfBlocks;
    if (blocks < 1) blocks = 1;
#define CALL(V) mf_fwd_kernel<V><<<(int)blocks, 256, 0, st>>>(user, item, rating, n, Wu, Wi, bu, bi, bias, pu, pi, \
                                                             colscale, add, loss_kind, pred, dpred, q_out, scratch)
    DMT_VEC_DISPATCH(H, CALL);
#undef CALL
    DMT_LAUNCH_CHECK();
    mf_finish_kernel<<<1, 256, 0, st>>>(scratch, (int)blocks, sums);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_mf_bwd_table(const int32_t* other, const float* W_other, const float* b_other, const float* p_side, int H,
                     const float* dpred, float grad_scale, const int32_t* perm, const int32_t* seg_key,
                     const int32_t* seg_off, const int32_t* n_seg, int64_t n_seg_max, const float* colscale, float* dW,
                     float* db, void* stream) {
    DMT_REQUIRE(n_seg_max >= 0 && dW, "dmt_mf_bwd_table: bad argument");
    if (n_seg_max == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n_seg_max + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define CALL(V) mf_bwd_table_kernel<V><<<(int)blocks, 256, 0, st>>>(other, W_other, b_other, p_side, dpred, grad_scale, \
                                                                   perm, seg_key, seg_off, n_seg, colscale, dW, db)
    DMT_VEC_DISPATCH(H, CALL);
#undef CALL
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_mf_bwd_side(const int32_t* idx, int64_t n, const float* W, const float* b, int H, const float* dpred,
                    float grad_scale, const float* colscale, float* d_p, void* stream) {
    DMT_REQUIRE(n >= 0 && d_p, "dmt_mf_bwd_side: bad argument");
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define CALL(V) mf_bwd_side_kernel<V><<<(int)blocks, 256, 0, st>>>(idx, n, W, b, dpred, grad_scale, colscale, d_p)
    DMT_VEC_DISPATCH(H, CALL);
#undef CALL
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_embed_fwd(const int32_t* idx, int64_t n, const float* W, const float* b, int H, float* out, int ld,
                  int col_off, void* stream) {
    DMT_REQUIRE(n >= 0 && out && ld >= col_off + H && (ld % 4) == 0 && (col_off % 4) == 0, "dmt_embed_fwd: bad argument");
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define CALL(V) embed_fwd_kernel<V><<<(int)blocks, 256, 0, st>>>(idx, n, W, b, out, ld, col_off)
    DMT_VEC_DISPATCH(H, CALL);
#undef CALL
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_embed_bwd(const float* dOut, int ld, int col_off, int H, const int32_t* perm, const int32_t* seg_key,
                  const int32_t* seg_off, const int32_t* n_seg, int64_t n_seg_max, float* dW, float* db, void* stream) {
    DMT_REQUIRE(n_seg_max >= 0 && dW && (ld % 4) == 0 && (col_off % 4) == 0, "dmt_embed_bwd: bad argument");
    if (n_seg_max == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n_seg_max + 7) / 8;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
#define CALL(V) embed_bwd_kernel<V><<<(int)blocks, 256, 0, st>>>(dOut, ld, col_off, perm, seg_key, seg_off, n_seg, dW, db)
    DMT_VEC_DISPATCH(H, CALL);
#undef CALL
    DMT_LAUNCH_CHECK();
    return 0;
}

int64_t dmt_weighted_colsum_scratch_floats(int width) { return (int64_t)kWcsBlocks * width; }

int dmt_weighted_colsum(const float* g, float scale, const float* Q, int64_t n, int width, int ld, float* out,
                        float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0 && width > 0 && ld >= width && out && scratch, "dmt_weighted_colsum: bad argument");
    cudaStream_t st = as_stream(stream);
    int nb = (int)(n < kWcsBlocks ? (n > 0 ? n : 1) : kWcsBlocks);
    weighted_colsum_stage1<<<nb, 256, 0, st>>>(g, scale, Q, n, width, ld, scratch);
    DMT_LAUNCH_CHECK();
    weighted_colsum_stage2<<<(width + 127) / 128, 128, 0, st>>>(scratch, nb, width, out);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_loss_fwd(const float* pred, const float* y, int64_t n, int loss_kind, float* dpred, float* sums,
                 float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0 && sums && scratch, "dmt_loss_fwd: bad argument");
    cudaStream_t st = as_stream(stream);
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > kMfBlocks) blocks = kMfBlocks;
    if (blocks < 1) blocks = 1;
    loss_fwd_kernel<<<(int)blocks, 256, 0, st>>>(pred, y, n, loss_kind, dpred, scratch);
    DMT_LAUNCH_CHECK();
    mf_finish_kernel<<<1, 256, 0, st>>>(scratch, (int)blocks, sums);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
