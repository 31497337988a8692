// MTAL coordinator kernels: pseudo-residual, weighted combination of organization outputs, the fused
// loss+gradient of the assisted-learning-rate / assistance-weight fit, and the round-0 mean predictor.
// All are HBM-streaming passes: one coalesced read of each operand, one write, fp32.
#include "common.cuh"

namespace dmt {

static thread_local char g_err[256] = "";
void set_error(const char* msg) {
    int i = 0;
    for (; msg[i] && i < 255; ++i) g_err[i] = msg[i];
    g_err[i] = 0;
}

// number of OUR kernels launched (captured launches are added when their graph is launched; engine.cu)
static long long g_launches = 0;
void count_launch(long long n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------- residual (reference src/assist.py:45-58)
__global__ void __launch_bounds__(256) residual_kernel(const float* __restrict__ F, const float* __restrict__ y,
                                                       float* __restrict__ r, int64_t n, int kind, float clamp) {
    int64_t n4 = n >> 2;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 f = ld4(F + 4 * i), t = ld4(y + 4 * i), o;
        float* fo = &o.x;
        const float* ff = &f.x;
        const float* tt = &t.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float g = loss_grad(kind, ff[k], tt[k]);
            if (clamp > 0.f) g = fminf(fmaxf(g, -clamp), clamp);
            fo[k] = -g;
        }
        st4(r + 4 * i, o);
    }
    // tail
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float g = loss_grad(kind, F[i], y[i]);
        if (clamp > 0.f) g = fminf(fmaxf(g, -clamp), clamp);
        r[i] = -g;
    }
}

// ---------------------------------------------------------------- combine (reference src/assist.py:131-176)
// One pass over the global CSR for all owners. S (K x K softmax rows) is staged in shared memory; with cold start
// (S_cold != NULL) a second K x K matrix holds softmax(w[1:]) per owner (slot 0 = 0): it is applied wherever
// organization 0's output is NaN, i.e. for the aligned rows organization 0 never saw (src/models/assist.py:28-34).
// org_row (may be NULL = identity) maps an organization id to its row of O (rank-blocked layouts, dist.py).
__global__ void __launch_bounds__(256) combine_kernel(const float* __restrict__ F_old, const float* __restrict__ O,
                                                      const int32_t* __restrict__ col,
                                                      const int32_t* __restrict__ owner,
                                                      const float* __restrict__ rate_col, const float* __restrict__ S,
                                                      const int64_t* __restrict__ match_end, float* __restrict__ F_new,
                                                      int64_t nnz, int K, const int32_t* __restrict__ org_row,
                                                      const float* __restrict__ S_cold) {
    extern __shared__ float sS[];  // K*K (+ K*K cold) + K row ids
    float* sC = sS + K * K;
    int* sRow = reinterpret_cast<int*>(sS + (S_cold ? 2 : 1) * K * K);
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
        sS[i] = S[i];
        if (S_cold) sC[i] = S_cold[i];
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x) sRow[i] = org_row ? org_row[i] : i;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) {
        int c = col[p];
        int ow = owner[c];
        const float* s = sS + ow * K;
        float q = 0.f;
        if (match_end == nullptr || p < match_end[ow]) {
            float o0 = O[(int64_t)sRow[0] * nnz + p];
            if (S_cold != nullptr && isnan(o0)) {
                const float* sc = sC + ow * K;
                for (int j = 1; j < K; ++j) q += O[(int64_t)sRow[j] * nnz + p] * sc[j];
            } else {
                q = o0 * s[0];
#pragma unroll 4
                for (int j = 1; j < K; ++j) q += O[(int64_t)sRow[j] * nnz + p] * s[j];
            }
        } else {
            // unmatched entry: the owner's own output fills every slot (reference src/assist.py:98-103)
            float own = O[(int64_t)sRow[ow] * nnz + p];
            for (int j = 0; j < K; ++j) q += own * s[j];
        }
        F_new[p] = F_old[p] + rate_col[c] * q;
    }
}

__global__ void __launch_bounds__(256) gather_view_kernel(const float* __restrict__ F_old, const float* __restrict__ y,
                                                          const float* __restrict__ O, const int32_t* __restrict__ pos,
                                                          const int32_t* __restrict__ rank, int64_t nnz, int64_t n,
                                                          int K, int owner, int64_t n_match, float* __restrict__ h,
                                                          float* __restrict__ t, float* __restrict__ V,
                                                          const int32_t* __restrict__ org_row) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        int64_t p = pos[e];
        h[e] = F_old[p];
        t[e] = y[p];
        bool matched = rank[e] < n_match;
        float own = O[(int64_t)(org_row ? org_row[owner] : owner) * nnz + p];
        for (int j = 0; j < K; ++j)
            V[(int64_t)j * n + e] = matched ? O[(int64_t)(org_row ? org_row[j] : j) * nnz + p] : own;
    }
}

// ---------------------------------------------------------------- models.Assist as a differentiable module
// forward of src/models/assist.py:25-37 over n entries: out is [n x K] addressed as out[e*se + j*sj] (so both the
// reference's row-major [n, K] stack and an organization-major [K, n] view fit); entries whose slot 0 is NaN (cold
// start) combine slots 1.. with softmax(w[1:]). q[e] keeps the weighted sum for the backward pass.
constexpr int kRowsKMax = 64;

__device__ __forceinline__ void softmax_pair(const float* __restrict__ w, int K, float* s, float* sc) {
    float mx = -INFINITY, mc = -INFINITY;
    for (int j = 0; j < K; ++j) {
        mx = fmaxf(mx, w[j]);
        if (j) mc = fmaxf(mc, w[j]);
    }
    float z = 0.f, zc = 0.f;
    for (int j = 0; j < K; ++j) {
        s[j] = expf(w[j] - mx);
        z += s[j];
        sc[j] = j ? expf(w[j] - mc) : 0.f;
        zc += sc[j];
    }
    for (int j = 0; j < K; ++j) {
        s[j] /= z;
        if (j) sc[j] /= zc;
    }
}

__global__ void __launch_bounds__(256) assist_rows_fwd_kernel(const float* __restrict__ out, int64_t se, int64_t sj,
                                                              const float* __restrict__ h,
                                                              const int32_t* __restrict__ idx,
                                                              const float* __restrict__ rate,
                                                              const float* __restrict__ w, int64_t n, int K,
                                                              float* __restrict__ tgt, float* __restrict__ q_out) {
    __shared__ float s[kRowsKMax], sc[kRowsKMax];
    if (threadIdx.x == 0) softmax_pair(w, K, s, sc);
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float* o = out + e * se;
        float o0 = o[0], q = 0.f;
        if (isnan(o0)) {
            for (int j = 1; j < K; ++j) q += o[j * sj] * sc[j];
        } else {
            q = o0 * s[0];
            for (int j = 1; j < K; ++j) q += o[j * sj] * s[j];
        }
        tgt[e] = h[e] + rate[idx[e]] * q;
        if (q_out) q_out[e] = q;
    }
}

// d_rate[seg_key[s]] = sum over the segment's entries of delta*q (entries sorted by idx: perm / seg_* of
// dmt_sort_segments); one warp per segment, lanes summed in a fixed order -> reproducible.
__global__ void __launch_bounds__(256) assist_rows_rate_grad_kernel(const int32_t* __restrict__ perm,
                                                                    const int32_t* __restrict__ seg_key,
                                                                    const int32_t* __restrict__ seg_off,
                                                                    const int32_t* __restrict__ n_seg,
                                                                    const float* __restrict__ delta,
                                                                    const float* __restrict__ q,
                                                                    float* __restrict__ d_rate) {
    int lane = threadIdx.x & 31;
    int warps = gridDim.x * (blockDim.x >> 5);
    int ns = n_seg[0];
    for (int sgm = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sgm < ns; sgm += warps) {
        int e0 = seg_off[sgm], e1 = seg_off[sgm + 1];
        float acc = 0.f;
        for (int e = e0 + lane; e < e1; e += 32) {
            int src = perm[e];
            acc += delta[src] * q[src];
        }
        acc = warp_sum(acc);
        if (lane == 0) d_rate[seg_key[sgm]] = acc;
    }
}

// block partials of d_s (warm entries) and d_sc (cold entries): scratch[b][2K]
__global__ void __launch_bounds__(256) assist_rows_w_grad_kernel(const float* __restrict__ out, int64_t se, int64_t sj,
                                                                 const int32_t* __restrict__ idx,
                                                                 const float* __restrict__ rate,
                                                                 const float* __restrict__ delta, int64_t n, int K,
                                                                 float* __restrict__ scratch) {
    __shared__ float sh[32];
    __shared__ float s_acc[8][2 * kRowsKMax];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = lane; j < 2 * K; j += 32) s_acc[wid][j] = 0.f;
    __syncwarp();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t n_round = (n + stride - 1) / stride * stride;  // warp-uniform trip count (shuffles need all lanes)
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_round; e += stride) {
        bool ok = e < n;
        const float* o = out + (ok ? e : 0) * se;
        float o0 = ok ? o[0] : 0.f;
        float de = ok ? delta[e] * rate[idx[e]] : 0.f;
        bool cold = isnan(o0);
        for (int j = 0; j < K; ++j) {
            float v = (j == 0) ? o0 : (ok ? o[j * sj] : 0.f);
            float warm = (!cold) ? de * v : 0.f;
            float cld = (cold && j > 0) ? de * v : 0.f;
            warm = warp_sum(warm);
            cld = warp_sum(cld);
            if (lane == 0) {
                s_acc[wid][j] += warm;
                s_acc[wid][K + j] += cld;
            }
        }
    }
    __syncthreads();
    (void)sh;
    for (int j = threadIdx.x; j < 2 * K; j += blockDim.x) {
        float v = 0.f;
        for (int i = 0; i < 8; ++i) v += s_acc[i][j];
        scratch[(int64_t)blockIdx.x * 2 * K + j] = v;
    }
}

__global__ void assist_rows_w_finish_kernel(const float* __restrict__ scratch, const float* __restrict__ w, int K,
                                            int nb, float* __restrict__ d_w) {
    __shared__ float s[kRowsKMax], sc[kRowsKMax], ds[2 * kRowsKMax];
    for (int j = threadIdx.x; j < 2 * K; j += blockDim.x) {
        float v = 0.f;
        for (int i = 0; i < nb; ++i) v += scratch[(int64_t)i * 2 * K + j];
        ds[j] = v;
    }
    if (threadIdx.x == 0) softmax_pair(w, K, s, sc);
    __syncthreads();
    if (threadIdx.x == 0) {
        float dot = 0.f, dotc = 0.f;
        for (int j = 0; j < K; ++j) {
            dot += s[j] * ds[j];
            dotc += sc[j] * ds[K + j];
        }
        for (int j = 0; j < K; ++j) d_w[j] = s[j] * (ds[j] - dot) + sc[j] * (ds[K + j] - dotc);
    }
}

// ---------------------------------------------------------------- L-BFGS on the device
// torch.optim.LBFGS(lr, max_iter = 20, max_eval = 25, tolerance_grad = 1e-7, tolerance_change = 1e-9,
// history_size = 100, no line search) as the reference runs it on models.assist (src/utils.py:255-256,
// src/assist.py:118-129), restated as ONE single-block kernel per inner iteration so that a whole fit is a chain of
// launches with no host round trip: the host enqueues, per optimizer.step(): closure, begin (it = 0), then for
// it = 1..max_iter: {check of iteration it-1 + direction / step of iteration it} and, while it < max_iter, the
// closure at the moved point, which skips itself (need_eval) once this step() call has met a stopping rule.
// Vectors live in global memory; reductions are block sums (one warp when P <= 32, e.g. the K assistance weights).
constexpr int kLbfgsStateFloats = 16;
struct LbfgsState {
    float loss, prev_loss, t, H_diag;
    int n_iter;     // state['n_iter'] of torch: counts over all step() calls
    int num_old, head;
    int step_done;  // this step() call has returned / broken out of its loop
    int need_eval;  // the closure after this kernel has to run (read by the closure kernels)
    int cur_evals;
    int pad[6];
};
static_assert(sizeof(LbfgsState) == kLbfgsStateFloats * sizeof(float), "state size");

__device__ __forceinline__ float block_max(float v, float* sh) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (lane < nw) ? sh[lane] : 0.f;  // all operands are magnitudes (>= 0)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    return r;
}

__global__ void lbfgs_step_kernel(LbfgsState* __restrict__ st, const float* __restrict__ loss_new,
                                  float* __restrict__ x, const float* __restrict__ g, float* __restrict__ d,
                                  float* __restrict__ prev_g, float* __restrict__ q, float* __restrict__ ty,
                                  float* __restrict__ ts, float* __restrict__ Y, float* __restrict__ S,
                                  float* __restrict__ ro, float* __restrict__ al, int P, int it, int max_iter, float lr,
                                  int hist) {
    __shared__ float sh[32];
    const float tol_grad = 1e-7f, tol_change = 1e-9f;
    const int max_eval = max_iter * 5 / 4;
    const int tid = threadIdx.x, nt = blockDim.x;
    LbfgsState s = *st;  // every thread holds the same copy; thread 0 writes it back
    __syncthreads();
    auto dot = [&](const float* a, const float* b) {
        float v = 0.f;
        for (int i = tid; i < P; i += nt) v += a[i] * b[i];
        return block_sum(v, sh);
    };
    auto absmax = [&](const float* a, float scale) {
        float v = 0.f;
        for (int i = tid; i < P; i += nt) v = fmaxf(v, fabsf(a[i] * scale));
        return block_max(v, sh);
    };
    auto finish = [&]() {
        if (tid == 0) *st = s;
    };
    if (it == 0) {  // after the first closure of a step() call
        s.loss = *loss_new;
        s.cur_evals = 1;
        s.need_eval = 0;
        s.step_done = absmax(g, 1.f) <= tol_grad ? 1 : 0;
        finish();
        return;
    }
    if (s.step_done) {
        s.need_eval = 0;
        finish();
        return;
    }
    if (s.need_eval) {  // iteration it-1 re-evaluated the closure: its stopping rules, in torch's order
        s.loss = *loss_new;
        s.cur_evals += 1;
        const bool opt_cond = absmax(g, 1.f) <= tol_grad;
        const float step_max = absmax(d, s.t);
        if (s.cur_evals >= max_eval || opt_cond || step_max <= tol_change || fabsf(s.loss - s.prev_loss) < tol_change) {
            s.step_done = 1;
            s.need_eval = 0;
            finish();
            return;
        }
    }
    // ---- iteration `it`: direction
    s.n_iter += 1;
    if (s.n_iter == 1) {
        for (int i = tid; i < P; i += nt) d[i] = -g[i];
        s.num_old = 0;
        s.head = 0;
        s.H_diag = 1.f;
    } else {
        for (int i = tid; i < P; i += nt) {
            ty[i] = g[i] - prev_g[i];
            ts[i] = d[i] * s.t;
        }
        __syncthreads();
        const float ys = dot(ty, ts);
        if (ys > 1e-10f) {
            if (s.num_old == hist) {  // old_dirs.pop(0)
                s.head = (s.head + 1) % hist;
                s.num_old -= 1;
            }
            const int slot = (s.head + s.num_old) % hist;
            for (int i = tid; i < P; i += nt) {
                Y[(int64_t)slot * P + i] = ty[i];
                S[(int64_t)slot * P + i] = ts[i];
            }
            if (tid == 0) ro[slot] = 1.f / ys;
            s.num_old += 1;
            __syncthreads();
            s.H_diag = ys / dot(ty, ty);
        }
        for (int i = tid; i < P; i += nt) q[i] = -g[i];
        __syncthreads();
        for (int j = s.num_old - 1; j >= 0; --j) {
            const int slot = (s.head + j) % hist;
            const float a = dot(S + (int64_t)slot * P, q) * ro[slot];
            if (tid == 0) al[slot] = a;
            for (int i = tid; i < P; i += nt) q[i] -= a * Y[(int64_t)slot * P + i];
            __syncthreads();
        }
        for (int i = tid; i < P; i += nt) d[i] = q[i] * s.H_diag;
        __syncthreads();
        for (int j = 0; j < s.num_old; ++j) {
            const int slot = (s.head + j) % hist;
            const float be = dot(Y + (int64_t)slot * P, d) * ro[slot];
            const float c = al[slot] - be;
            for (int i = tid; i < P; i += nt) d[i] += c * S[(int64_t)slot * P + i];
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < P; i += nt) prev_g[i] = g[i];
    s.prev_loss = s.loss;
    // ---- step length
    if (s.n_iter == 1) {
        float v = 0.f;
        for (int i = tid; i < P; i += nt) v += fabsf(g[i]);
        v = block_sum(v, sh);
        s.t = fminf(1.f, 1.f / v) * lr;
    } else {
        s.t = lr;
    }
    const float gtd = dot(g, d);
    if (gtd > -tol_change) {  // directional derivative below tolerance: break before moving
        s.step_done = 1;
        s.need_eval = 0;
        finish();
        return;
    }
    for (int i = tid; i < P; i += nt) x[i] += s.t * d[i];
    if (it != max_iter) {
        s.need_eval = 1;
    } else {
        s.need_eval = 0;
        s.step_done = 1;
    }
    finish();
}

// ---------------------------------------------------------------- fused loss + grad of models.Assist
// scratch layout: [0] loss partials (NB), then d_s partials (NB x K). One warp per owned column (segment) so the
// per-column rate gradient is a warp-segmented reduction without atomics; the K-vector d_s is reduced per block
// and finished by a second tiny kernel (deterministic two-stage).
constexpr int kAssistBlocks = kNumSMs * 2;
constexpr int kAssistKMax = 64;

__global__ void __launch_bounds__(256) assist_loss_grad_kernel(const float* __restrict__ h, const float* __restrict__ t,
                                                               const float* __restrict__ V,
                                                               const int32_t* __restrict__ seg_off,
                                                               const float* __restrict__ rate,
                                                               const float* __restrict__ w, int64_t n, int n_rate, int K,
                                                               int kind, float* __restrict__ d_rate,
                                                               float* __restrict__ scratch,
                                                               const int* __restrict__ run_flag) {
    __shared__ float s_soft[kAssistKMax];
    __shared__ float s_ds[8][kAssistKMax];
    __shared__ float s_loss[8];
    if (run_flag != nullptr && *run_flag == 0) return;  // device-side L-BFGS: this evaluation is not needed
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        float mx = -INFINITY;
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, w[j]);
        float z = 0.f;
        for (int j = 0; j < K; ++j) {
            s_soft[j] = expf(w[j] - mx);
            z += s_soft[j];
        }
        for (int j = 0; j < K; ++j) s_soft[j] /= z;
    }
    for (int j = lane; j < K; j += 32) s_ds[wid][j] = 0.f;
    __syncthreads();
    float inv_n = 1.f / (float)n;
    float loss_acc = 0.f;
    int warps_total = gridDim.x * 8;
    for (int c = blockIdx.x * 8 + wid; c < n_rate; c += warps_total) {
        int e0 = seg_off[c], e1 = seg_off[c + 1];
        float eta = rate[c];
        float dr = 0.f;
        for (int eb = e0; eb < e1; eb += 32) {  // warp-uniform trip count: the shuffles below need all lanes
            int e = eb + lane;
            bool ok = e < e1;
            float q = 0.f, de = 0.f;
            if (ok) {
                for (int j = 0; j < K; ++j) q += V[(int64_t)j * n + e] * s_soft[j];
                float o = h[e] + eta * q;
                float y = t[e];
                loss_acc += loss_value(kind, o, y);
                float delta = loss_grad(kind, o, y) * inv_n;
                dr += delta * q;
                de = delta * eta;
            }
            for (int j = 0; j < K; ++j) {
                float v = ok ? de * V[(int64_t)j * n + e] : 0.f;
                v = warp_sum(v);
                if (lane == 0) s_ds[wid][j] += v;
            }
        }
        dr = warp_sum(dr);
        if (lane == 0) d_rate[c] = dr;
    }
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) s_loss[wid] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        for (int i = 0; i < 8; ++i) l += s_loss[i];
        scratch[blockIdx.x] = l * inv_n;
    }
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        float v = 0.f;
        for (int i = 0; i < 8; ++i) v += s_ds[i][j];
        scratch[gridDim.x + (int64_t)blockIdx.x * K + j] = v;
    }
}

__global__ void assist_finish_kernel(const float* __restrict__ scratch, const float* __restrict__ w, int K, int nb,
                                     float* __restrict__ out_loss, float* __restrict__ d_w,
                                     const int* __restrict__ run_flag) {
    __shared__ float sh[32];
    __shared__ float s_soft[kAssistKMax];
    __shared__ float s_ds[kAssistKMax];
    if (run_flag != nullptr && *run_flag == 0) return;
    float l = 0.f;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) l += scratch[i];
    l = block_sum(l, sh);
    if (threadIdx.x == 0) out_loss[0] = l;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        float v = 0.f;
        for (int i = 0; i < nb; ++i) v += scratch[nb + (int64_t)i * K + j];
        s_ds[j] = v;
    }
    if (threadIdx.x == 0) {
        float mx = -INFINITY;
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, w[j]);
        float z = 0.f;
        for (int j = 0; j < K; ++j) {
            s_soft[j] = expf(w[j] - mx);
            z += s_soft[j];
        }
        for (int j = 0; j < K; ++j) s_soft[j] /= z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // d/dw of softmax: dw_j = s_j (ds_j - sum_k s_k ds_k)
        float dot = 0.f;
        for (int j = 0; j < K; ++j) dot += s_soft[j] * s_ds[j];
        for (int j = 0; j < K; ++j) d_w[j] = s_soft[j] * (s_ds[j] - dot);
    }
}

// ---------------------------------------------------------------- models.Base (reference src/models/base.py:22-60)
__global__ void __launch_bounds__(256) base_fit_kernel(const int32_t* __restrict__ idx, const float* __restrict__ rating,
                                                       int64_t n, float* base, float* count) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // ratings are small integers ({1..5} or {0,1}): fp32 sums are exact, so the atomic order cannot matter
        atomicAdd(base + idx[i], rating[i]);
        atomicAdd(count + idx[i], 1.f);
    }
}

// mean of the seen means -> scratch[0]; single block (n_cols is the number of columns of one organization)
__global__ void base_fill_kernel(const float* __restrict__ base, const float* __restrict__ count, int n_cols,
                                 float* scratch) {
    __shared__ float sh[32];
    float s = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < n_cols; i += blockDim.x) {
        if (count[i] != 0.f) {
            s += base[i] / count[i];
            c += 1.f;
        }
    }
    s = block_sum(s, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) scratch[0] = s / c;
}

__global__ void __launch_bounds__(256) base_predict_kernel(const float* __restrict__ base,
                                                           const float* __restrict__ count,
                                                           const int32_t* __restrict__ tidx, int64_t n, int implicit,
                                                           float implicit_count, const float* __restrict__ fill,
                                                           float* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int c = tidx[i];
        if (implicit) {
            out[i] = base[c] / implicit_count;
        } else {
            float cnt = count[c];
            out[i] = (cnt == 0.f) ? fill[0] : base[c] / (cnt + 1e-10f);
        }
    }
}

static inline int stream_grid(int64_t n, int per_thread = 4) {
    int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

const char* dmt_last_error(void) { return g_err; }
int dmt_version(void) { return 100; }
int64_t dmt_launch_count(void) { return (int64_t)launch_count(); }

int dmt_check_device(void) {
    int dev = 0;
    DMT_CUDA(cudaGetDevice(&dev));
    int major = 0;
    DMT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("libdmt_b200 is built for sm_100a only");
        return DMT_E_ARCH;
    }
    return 0;
}

int dmt_residual(const float* F, const float* y, float* r, int64_t n, int loss_kind, float clamp, void* stream) {
    DMT_REQUIRE(n >= 0 && (loss_kind == DMT_LOSS_MSE || loss_kind == DMT_LOSS_BCE), "dmt_residual: bad argument");
    if (n == 0) return 0;
    residual_kernel<<<stream_grid(n, 8), 256, 0, as_stream(stream)>>>(F, y, r, n, loss_kind, clamp);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_assist_combine(const float* F_old, const float* O, const int32_t* col, const int32_t* owner,
                       const float* rate_col, const float* S, const int64_t* match_end, float* F_new, int64_t nnz,
                       int K, const int32_t* org_row, const float* S_cold, void* stream) {
    DMT_REQUIRE(nnz >= 0 && K >= 1 && K <= 128, "dmt_assist_combine: need 1 <= K <= 128");
    if (nnz == 0) return 0;
    size_t smem = ((S_cold ? 2 : 1) * (size_t)K * K + K) * sizeof(float);
    if (smem > 48 * 1024) {
        DMT_REQUIRE(smem <= 200 * 1024, "dmt_assist_combine: K too large for the shared-memory weight tables");
        DMT_CUDA(cudaFuncSetAttribute(combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    combine_kernel<<<stream_grid(nnz, 2), 256, smem, as_stream(stream)>>>(F_old, O, col, owner, rate_col, S, match_end,
                                                                          F_new, nnz, K, org_row, S_cold);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_assist_gather_view(const float* F_old, const float* y, const float* O, const int32_t* pos,
                           const int32_t* rank, int64_t nnz, int64_t n, int K, int owner, int64_t n_match, float* h,
                           float* t, float* V, const int32_t* org_row, void* stream) {
    DMT_REQUIRE(n >= 0 && K >= 1 && owner >= 0 && owner < K, "dmt_assist_gather_view: bad argument");
    if (n == 0) return 0;
    gather_view_kernel<<<stream_grid(n, 1), 256, 0, as_stream(stream)>>>(F_old, y, O, pos, rank, nnz, n, K, owner,
                                                                         n_match, h, t, V, org_row);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_assist_rows_fwd(const float* out, int64_t stride_e, int64_t stride_j, const float* history, const int32_t* idx,
                        const float* rate, const float* w, int64_t n, int K, float* target, float* q, void* stream) {
    DMT_REQUIRE(n >= 0 && K >= 1 && K <= kRowsKMax, "dmt_assist_rows_fwd: need 1 <= K <= 64");
    if (n == 0) return 0;
    assist_rows_fwd_kernel<<<stream_grid(n, 2), 256, 0, as_stream(stream)>>>(out, stride_e, stride_j, history, idx, rate,
                                                                             w, n, K, target, q);
    DMT_LAUNCH_CHECK();
    return 0;
}

int64_t dmt_assist_rows_scratch_floats(int K) { return (int64_t)kAssistBlocks * 2 * K; }

int dmt_assist_rows_bwd(const float* out, int64_t stride_e, int64_t stride_j, const int32_t* idx, const float* rate,
                        const float* w, const float* q, const float* delta, const int32_t* perm, const int32_t* seg_key,
                        const int32_t* seg_off, const int32_t* n_seg, int64_t n, int K, int n_rate, float* d_rate,
                        float* d_w, float* scratch, void* stream) {
    DMT_REQUIRE(n > 0 && K >= 1 && K <= kRowsKMax && n_rate > 0, "dmt_assist_rows_bwd: need n>0 and 1 <= K <= 64");
    if (d_rate != nullptr) {
        DMT_CUDA(cudaMemsetAsync(d_rate, 0, sizeof(float) * (size_t)n_rate, as_stream(stream)));
        int64_t segs = n < n_rate ? n : n_rate;
        int nbr = (int)((segs + 7) / 8);
        if (nbr > kAssistBlocks) nbr = kAssistBlocks;
        assist_rows_rate_grad_kernel<<<nbr, 256, 0, as_stream(stream)>>>(perm, seg_key, seg_off, n_seg, delta, q, d_rate);
        DMT_LAUNCH_CHECK();
    }
    if (d_w != nullptr) {
        int nb = stream_grid(n, 4);
        if (nb > kAssistBlocks) nb = kAssistBlocks;
        assist_rows_w_grad_kernel<<<nb, 256, 0, as_stream(stream)>>>(out, stride_e, stride_j, idx, rate, delta, n, K,
                                                                     scratch);
        DMT_LAUNCH_CHECK();
        assist_rows_w_finish_kernel<<<1, 128, 0, as_stream(stream)>>>(scratch, w, K, nb, d_w);
        DMT_LAUNCH_CHECK();
    }
    return 0;
}

int64_t dmt_assist_scratch_floats(int K) { return (int64_t)kAssistBlocks * (1 + K); }

int dmt_assist_loss_grad(const float* h, const float* t, const float* V, const int32_t* seg_off, const float* rate,
                         const float* w, int64_t n, int n_rate, int K, int loss_kind, float* out_loss, float* d_rate,
                         float* d_w, float* scratch, void* stream) {
    DMT_REQUIRE(n > 0 && n_rate > 0 && K >= 1 && K <= kAssistKMax, "dmt_assist_loss_grad: need n>0 and 1 <= K <= 64");
    int nb = (n_rate + 7) / 8;
    if (nb > kAssistBlocks) nb = kAssistBlocks;
    assist_loss_grad_kernel<<<nb, 256, 0, as_stream(stream)>>>(h, t, V, seg_off, rate, w, n, n_rate, K, loss_kind,
                                                              d_rate, scratch, nullptr);
    DMT_LAUNCH_CHECK();
    assist_finish_kernel<<<1, 256, 0, as_stream(stream)>>>(scratch, w, K, nb, out_loss, d_w, nullptr);
    DMT_LAUNCH_CHECK();
    return 0;
}

// Workspace of one device-side fit (floats): state | loss_new | grads[n_rate + K] | d, prev_g, q, ty, ts [5 x P] |
// ro, al [2 x H] | old_dirs, old_stps [2 x H x P], with P = n_rate + K (the upper bound of the optimised slice).
int64_t dmt_assist_fit_work_floats(int n_rate, int K, int history) {
    const int64_t Pm = (int64_t)n_rate + K;
    return kLbfgsStateFloats + 1 + Pm + 5 * Pm + 2 * (int64_t)history + 2 * (int64_t)history * Pm;
}

int dmt_assist_fit(const float* h, const float* t, const float* V, const int32_t* seg_off, int64_t n, int n_rate, int K,
                   int loss_kind, float* params, int ar_optim, int aw_optim, float lr, int steps, int max_iter,
                   int history, float* work, float* scratch, void* stream) {
    DMT_REQUIRE(n > 0 && n_rate > 0 && K >= 1 && K <= kAssistKMax && params && work && scratch,
                "dmt_assist_fit: need n>0 and 1 <= K <= 64");
    DMT_REQUIRE((ar_optim || aw_optim) && steps > 0 && max_iter > 0 && history > 0, "dmt_assist_fit: nothing to fit");
    cudaStream_t st = as_stream(stream);
    const int64_t Pm = (int64_t)n_rate + K;
    const int off = ar_optim ? 0 : n_rate;
    const int P = (ar_optim ? n_rate : 0) + (aw_optim ? K : 0);
    LbfgsState* state = reinterpret_cast<LbfgsState*>(work);
    float* loss_new = work + kLbfgsStateFloats;
    float* grads = loss_new + 1;
    float* vec = grads + Pm;  // d | prev_g | q | ty | ts
    float* ro = vec + 5 * Pm;
    float* al = ro + history;
    float* Y = al + history;
    float* S = Y + (int64_t)history * Pm;
    DMT_CUDA(cudaMemsetAsync(work, 0, sizeof(float) * (kLbfgsStateFloats + 1), st));
    int nb = (n_rate + 7) / 8;
    if (nb > kAssistBlocks) nb = kAssistBlocks;
    const int threads = P <= 32 ? 32 : 256;
    const float* rate = params;
    const float* w = params + n_rate;
    auto eval = [&](const int* flag) -> int {
        assist_loss_grad_kernel<<<nb, 256, 0, st>>>(h, t, V, seg_off, rate, w, n, n_rate, K, loss_kind, grads, scratch,
                                                    flag);
        DMT_LAUNCH_CHECK();
        assist_finish_kernel<<<1, 256, 0, st>>>(scratch, w, K, nb, loss_new, grads + n_rate, flag);
        DMT_LAUNCH_CHECK();
        return 0;
    };
    int rc;
    for (int s_ = 0; s_ < steps; ++s_) {  // optimizer.step(closure) x cfg['assist']['num_epochs'] (src/assist.py:118-129)
        if ((rc = eval(nullptr))) return rc;
        for (int it = 0; it <= max_iter; ++it) {
            lbfgs_step_kernel<<<1, threads, 0, st>>>(state, loss_new, params + off, grads + off, vec, vec + Pm,
                                                     vec + 2 * Pm, vec + 3 * Pm, vec + 4 * Pm, Y, S, ro, al, P, it,
                                                     max_iter, lr, history);
            DMT_LAUNCH_CHECK();
            if (it >= 1 && it < max_iter)
                if ((rc = eval(&state->need_eval))) return rc;
        }
    }
    return 0;
}

int dmt_base_fit(const int32_t* idx, const float* rating, int64_t n, float* base, float* count, void* stream) {
    DMT_REQUIRE(n >= 0, "dmt_base_fit: bad n");
    if (n == 0) return 0;
    base_fit_kernel<<<stream_grid(n, 2), 256, 0, as_stream(stream)>>>(idx, rating, n, base, count);
    DMT_LAUNCH_CHECK();
    return 0;
}

int dmt_base_predict(const float* base, const float* count, int32_t n_cols, const int32_t* target_idx, int64_t n,
                     int implicit, float implicit_count, float* out, float* scratch, void* stream) {
    DMT_REQUIRE(n >= 0 && n_cols > 0, "dmt_base_predict: bad argument");
    if (!implicit) {
        base_fill_kernel<<<1, 1024, 0, as_stream(stream)>>>(base, count, n_cols, scratch);
        DMT_LAUNCH_CHECK();
    }
    if (n == 0) return 0;
    base_predict_kernel<<<stream_grid(n, 2), 256, 0, as_stream(stream)>>>(base, count, target_idx, n, implicit,
                                                                          implicit_count, scratch, out);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
