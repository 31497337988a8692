// Evaluation of the global prediction on the device (reference src/train_recsys_assist.py:175-217 `test()`):
// the split is walked in blocks of `block_rows` rows (the reference's test batch); per block
//   Loss  = mean loss over the block's entries              (src/models/utils.py:7-14)
//   RMSE  = sqrt(mean squared error)                        (src/metrics/metrics.py:8-11)
//   NDCG  = mean over the block's rows WITH entries of DCG@k / IDCG@k, where the block is densified over its observed
//           columns: unobserved scores are -inf (ranked last) and unobserved gains 0 (src/metrics/metrics.py:63-84);
//           k = min(topk, #distinct columns of the block); 0/0 and x/0 count as 0.
// and the caller forms the entry-weighted mean over blocks (src/logger.py:35-55). One CTA per block, one warp per row,
// fixed summation orders (no atomics): bit-reproducible. Segmented top-k instead of the reference's dense
// [rows x columns] scatter: k rounds of a warp arg-max over the row's entries in (value desc, position asc) order.
#include "kernels.cuh"

namespace dmt {

__device__ __forceinline__ bool ranks_before(float v, int i, float pv, int pi) {  // (v, i) strictly after (pv, pi)?
    return v < pv || (v == pv && i > pi);
}

// DCG of the row's top-k entries by `key` with gains `gain` (key == gain gives the ideal DCG)
__device__ __forceinline__ float row_dcg(const float* __restrict__ key, const float* __restrict__ gain, int e0, int e1,
                                         int k, int lane) {
    float prev_v = INFINITY;
    int prev_i = -1;
    float dcg = 0.f;
    for (int r = 0; r < k; ++r) {
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
        for (int e = e0 + lane; e < e1; e += 32) {
            const float v = key[e];
            if (!(r == 0 || ranks_before(v, e, prev_v, prev_i))) continue;
            if (v > best_v || (v == best_v && e < best_i)) {
                best_v = v;
                best_i = e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) {
                best_v = ov;
                best_i = oi;
            }
        }
        if (best_i == 0x7fffffff) break;  // fewer than k entries: the rest ranks unobserved columns (gain 0)
        dcg += gain[best_i] / log2f((float)(r + 2));
        prev_v = best_v;
        prev_i = best_i;
    }
    return dcg;
}

__global__ void __launch_bounds__(256) eval_blocks_kernel(const int32_t* __restrict__ indptr,
                                                          const float* __restrict__ pred,
                                                          const float* __restrict__ target, int n_rows,
                                                          int block_rows, int loss_kind, int want_ndcg,
                                                          const int32_t* __restrict__ block_k,
                                                          float* __restrict__ out) {
    __shared__ float sh[3][8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r0 = b * block_rows, r1 = min(n_rows, r0 + block_rows);
    const int k = want_ndcg ? block_k[b] : 0;
    float loss = 0.f, sq = 0.f, q_sum = 0.f;
    for (int r = r0 + wid; r < r1; r += 8) {
        const int e0 = indptr[r], e1 = indptr[r + 1];
        if (e1 == e0) continue;
        float l = 0.f, s = 0.f;
        for (int e = e0 + lane; e < e1; e += 32) {
            const float o = pred[e], y = target[e];
            l += loss_value(loss_kind, o, y);
            s += (o - y) * (o - y);
        }
        loss += warp_sum(l);
        sq += warp_sum(s);
        if (want_ndcg) {
            const float dcg = row_dcg(pred, target, e0, e1, k, lane);
            const float idcg = row_dcg(target, target, e0, e1, k, lane);
            const float q = dcg / idcg;
            q_sum += (isnan(q) || isinf(q)) ? 0.f : q;
        }
    }
    if (lane == 0) {
        sh[0][wid] = loss;
        sh[1][wid] = sq;
        sh[2][wid] = q_sum;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        out[b * 3 + threadIdx.x] = t;
    }
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int dmt_eval_blocks(const int32_t* indptr, const float* pred, const float* target, int n_rows, int block_rows,
                    int loss_kind, int want_ndcg, const int32_t* block_k, float* out, void* stream) {
    DMT_REQUIRE(indptr && pred && target && out && n_rows >= 0 && block_rows > 0, "dmt_eval_blocks: bad argument");
    DMT_REQUIRE(!want_ndcg || block_k, "dmt_eval_blocks: NDCG needs the per-block k");
    const int n_blocks = (n_rows + block_rows - 1) / block_rows;
    if (n_blocks == 0) return 0;
    eval_blocks_kernel<<<n_blocks, 256, 0, as_stream(stream)>>>(indptr, pred, target, n_rows, block_rows, loss_kind,
                                                               want_ndcg, block_k, out);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
