// AAE decoder last layer as dense tensor-core GEMMs with sparse ends (reference src/models/ae.py:135-156):
//
//   D1  O  = A3 . W4^T            [B' x N]   epilogue: + b4, masked to the batch's target entries (CSR), loss,
//                                            g = dL/do written per target entry; nothing dense leaves the SM
//   D2  dZ3 = (G . W4) * (1-A3^2) [B' x H]   G = the sparse [B' x N] matrix of g, scattered from the CSR straight into
//                                            the swizzled shared-memory operand tile; split-K over the items
//   D3  dW4 = G^T . A3            [N x H]    G^T tiles scattered from the CSR the same way; db4 = row sums of G^T
//
// All three run on tcgen05 (kind::tf32, accumulators in TMEM) with the 3xTF32 split of umma.cuh, which keeps the
// fp32 parity bar (loss within 1e-5 relative). The dense products cost 2*B'*N*H flop each (0.95 GFLOP at ML1M shape)
// instead of 1 KB of L2 gather per target entry; at ML1M/Douban/Amazon density that is the cheaper form, and the
// engine keeps the gather/SDDMM form (ae.cu, segments.cu) for very sparse shapes.
// Needs ascending column indices inside every CSR row (the engine checks at ingestion).
#include "kernels.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

constexpr int kSStride = TN + 1;  // padded row stride of the accumulator tile in shared memory (conflict-free)
constexpr int align16(int x) { return (x + 15) / 16 * 16; }
// D1's area behind the operand tiles: the tail of the accumulator tile, row offsets/starts/output bases, reduction scratch
constexpr int kFwdExtra = align16(TM * kSStride * 4 - kOperandBytes + (TM + 1) * 4 + 2 * TM * 4 + 32 * 4);

struct DecTcArgs {
    const int32_t* rows;     // batch-row (epoch-wide index) -> row of the CSR
    const int32_t* indptr;
    const int32_t* indices;
    const float* target;     // may be null (predict)
    const float* A3;         // [B' x H], row 0 = first row of the batch
    const float* W4;         // [n_dec x H]
    const float* b4;
    int H, n_dec, loss_kind;
    const int32_t* n_targets;  // [batch] (engine) or [1]
    const int32_t* ent_off;    // epoch-wide batch-row -> offset in the batch-ordered entry space; null: CSR positions
    float* gout;               // dL/do per target entry (train) or null
    float* pred;               // o per target entry or null
    float* loss_part;          // [grid.y * grid.x] per-tile loss sums (train)
};

// ---------------------------------------------------------------- D1: forward + masked loss epilogue
__global__ void __launch_bounds__(kThreads) dec_fwd_tc_kernel(DecTcArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    if (m0 >= M) return;
    Ctx c = setup(smem_raw, kFwdExtra);
    const int t = threadIdx.x;
    // this row's target entries with a column inside the tile
    int e0 = 0, s = 0, cnt = 0, obase = 0;
    if (m0 + t < M) {
        const int u = p.rows[lo + m0 + t];
        e0 = p.indptr[u];
        const int e1 = p.indptr[u + 1];
        s = lower_bound_i32(p.indices, e0, e1, n0);
        cnt = lower_bound_i32(p.indices, s, e1, n0 + TN) - s;
        obase = p.ent_off ? p.ent_off[lo + m0 + t] - e0 : 0;
    }
    for (int k0 = 0; k0 < p.H; k0 += TK) {
        stage_kcontig(p.A3, p.H, m0, M, k0, p.H, c.A_hi, c.A_lo, passes);
        stage_kcontig(p.W4, p.H, n0, p.n_dec, k0, p.H, c.B_hi, c.B_lo, passes);
        issue(c, passes);
        wait(c);
    }
    // accumulator tile -> shared memory (aliases the operand tiles: every MMA has finished reading them)
    float* S = reinterpret_cast<float*>(c.base);
    int32_t* s_off = reinterpret_cast<int32_t*>(c.base + TM * kSStride * 4);  // [TM + 1]
    int32_t* s_start = s_off + TM + 1;                                       // [TM]
    int32_t* s_obase = s_start + TM;                                         // [TM]
    float* s_red = reinterpret_cast<float*>(s_obase + TM);                   // [32]
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
        float v[32];
        load_acc32(c, c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) S[t * kSStride + c0 + j] = v[j];
    }
    s_off[t] = cnt;  // counts first, scanned in place below
    s_start[t] = s;
    s_obase[t] = obase;
    __syncthreads();
    if (t < 32) {  // exclusive scan of the 128 counts: 4 per lane
        const int a0 = s_off[4 * t], a1 = s_off[4 * t + 1], a2 = s_off[4 * t + 2], a3 = s_off[4 * t + 3];
        const int sum4 = a0 + a1 + a2 + a3;
        int incl = sum4;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (t >= o) incl += n;
        }
        const int excl = incl - sum4;
        s_off[4 * t] = excl;
        s_off[4 * t + 1] = excl + a0;
        s_off[4 * t + 2] = excl + a0 + a1;
        s_off[4 * t + 3] = excl + a0 + a1 + a2;
        if (t == 31) s_off[TM] = incl;
    }
    __syncthreads();
    const int total = s_off[TM];
    const bool train = p.gout != nullptr;
    const float inv_n = train ? 1.f / (float)p.n_targets[br.row_off ? br.b : 0] : 0.f;
    float loss_acc = 0.f;
    for (int j = t; j < total; j += kThreads) {  // the tile's target entries, flattened: balanced over the threads
        int rl = 0, rh = TM;                     // row r with s_off[r] <= j < s_off[r + 1]
        while (rh - rl > 1) {
            const int mid = (rl + rh) >> 1;
            if (s_off[mid] <= j) rl = mid; else rh = mid;
        }
        const int e = s_start[rl] + (j - s_off[rl]);
        const int col = p.indices[e];
        const float o = S[rl * kSStride + (col - n0)] + p.b4[col];
        const int64_t pos = (int64_t)s_obase[rl] + e;
        if (p.pred) p.pred[pos] = o;
        if (train) {
            const float y = p.target[e];
            p.gout[pos] = loss_grad(p.loss_kind, o, y) * inv_n;
            loss_acc += loss_value(p.loss_kind, o, y);
        }
    }
    if (train) {
        loss_acc = block_sum(loss_acc, s_red);
        if (t == 0) p.loss_part[blockIdx.y * gridDim.x + blockIdx.x] = loss_acc;
    }
    teardown(c);
}

// ---------------------------------------------------------------- D2: dA3 partials = G . W4 over one K split
struct DecBwdAArgs {
    const int32_t* rows;
    const int32_t* indptr;
    const int32_t* indices;
    const float* gbuf;       // g per target entry
    const int32_t* ent_off;  // null: gbuf is at CSR positions
    const float* W4;
    int H, n_dec;
    int chunks_per_split;    // k-chunks (of 32 items) per grid.z slice
    float* part;             // [grid.z][part_rows x H]
    int part_rows;
};

__global__ void __launch_bounds__(kThreads) dec_bwd_a_tc_kernel(DecBwdAArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    if (m0 >= M) return;
    Ctx c = setup(smem_raw, 0);
    const int t = threadIdx.x;
    const int n_chunks = (p.n_dec + TK - 1) / TK;
    const int kc0 = blockIdx.z * p.chunks_per_split;
    const int kc1 = min(n_chunks, kc0 + p.chunks_per_split);
    int ptr = 0, e1 = 0;
    const float* grow = p.gbuf;
    if (m0 + t < M) {  // this thread walks row m0 + t of G through the split's item range
        const int u = p.rows[lo + m0 + t];
        const int e0 = p.indptr[u];
        e1 = p.indptr[u + 1];
        ptr = lower_bound_i32(p.indices, e0, e1, kc0 * TK);
        if (p.ent_off) grow = p.gbuf + ((int64_t)p.ent_off[lo + m0 + t] - e0);
    }
    for (int kc = kc0; kc < kc1; ++kc) {
        const int k0 = kc * TK;
        zero_tiles(c.A_hi, c.A_lo, passes);
        stage_transposed(p.W4, p.H, n0, p.H, k0, p.n_dec, c.B_hi, c.B_lo, passes);  // B(n = unit, k = item) = W4[item][unit]
        __syncthreads();  // canvas zeroed by all threads before any scatter
        while (ptr < e1) {
            const int col = p.indices[ptr];
            if (col >= k0 + TK) break;
            store_split1(c.A_hi, c.A_lo, tile_off(t, col - k0), grow[ptr], passes);
            ++ptr;
        }
        issue(c, passes);
        wait(c);
    }
    float* out = p.part + ((int64_t)blockIdx.z * p.part_rows + (m0 + t)) * p.H + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
        float v[32];
        load_acc32(c, c0, v);  // warp-collective: every lane takes part, rows past the batch just do not store
        if (m0 + t < M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) st4(out + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
    }
    teardown(c);
}

// dZ3 = (sum of the K-split partials, in split order) * (1 - A3^2); block per batch row. Also folds the per-tile loss
// sums of D1 into loss_rows (row 0 carries the batch's loss sum, the other rows 0) for adam_prepare.
__global__ void __launch_bounds__(256) dec_bwd_a_finish_kernel(const float* __restrict__ part, int splits, int part_rows,
                                                               const float* __restrict__ A3, int H, int tanh_deriv,
                                                               float* __restrict__ dZ3,
                                                               const float* __restrict__ loss_part, int tiles_x,
                                                               float* __restrict__ loss_rows, BatchRef br) {
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int j = blockIdx.x;
    if (j >= M) return;
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[((int64_t)z * part_rows + j) * H + h];
        const float a = A3[(int64_t)j * H + h];
        dZ3[(int64_t)j * H + h] = tanh_deriv ? s * (1.f - a * a) : s;
    }
    if (loss_rows != nullptr && threadIdx.x == 0) {
        float l = 0.f;
        if (j == 0) {
            const int n_tiles = ((M + TM - 1) / TM) * tiles_x;
            for (int i = 0; i < n_tiles; ++i) l += loss_part[i];
        }
        loss_rows[j] = l;
    }
}

// ---------------------------------------------------------------- D3: dW4 = G^T . A3, db4 = row sums of G^T
struct DecBwdWArgs {
    const int32_t* rows;
    const int32_t* indptr;
    const int32_t* indices;
    const float* gbuf;
    const int32_t* ent_off;
    const float* A3;
    int H, n_dec;
    float* dW4;  // [n_dec x H]
    float* db4;  // [n_dec] or null
    int rows_cap;  // capacity of the per-row index arrays in shared memory (>= batch rows)
};

__global__ void __launch_bounds__(kThreads) dec_bwd_w_tc_kernel(DecBwdWArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM /* item tile */, n0 = blockIdx.x * TN /* hidden units */;
    const int extra = align16(3 * p.rows_cap * 4);
    Ctx c = setup(smem_raw, extra);
    const int t = threadIdx.x;
    int32_t* r_start = reinterpret_cast<int32_t*>(c.base + kOperandBytes);  // first entry of the row inside the tile
    int32_t* r_end = r_start + p.rows_cap;
    int32_t* r_gofs = r_end + p.rows_cap;  // gbuf index of entry e is r_gofs + e
    for (int r = t; r < M; r += kThreads) {
        const int u = p.rows[lo + r];
        const int e0 = p.indptr[u], e1 = p.indptr[u + 1];
        const int s = lower_bound_i32(p.indices, e0, e1, m0);
        r_start[r] = s;
        r_end[r] = lower_bound_i32(p.indices, s, e1, m0 + TM);
        r_gofs[r] = p.ent_off ? p.ent_off[lo + r] - e0 : 0;
    }
    __syncthreads();
    const bool want_db = p.db4 != nullptr && blockIdx.x == 0;
    float db_acc = 0.f;
    for (int k0 = 0; k0 < M; k0 += TK) {
        zero_tiles(c.A_hi, c.A_lo, passes);
        stage_transposed(p.A3, p.H, n0, p.H, k0, M, c.B_hi, c.B_lo, passes);  // B(n = unit, k = batch row) = A3[row][unit]
        __syncthreads();
        {   // four threads per batch row of the chunk scatter its in-tile entries: tile(item - m0, row - k0) = g
            const int r = k0 + (t >> 2);
            if (r < M) {
                const int e_end = r_end[r], gofs = r_gofs[r];
                for (int e = r_start[r] + (t & 3); e < e_end; e += 4)
                    store_split1(c.A_hi, c.A_lo, tile_off(p.indices[e] - m0, r - k0), p.gbuf[(int64_t)gofs + e], passes);
            }
        }
        issue(c, passes);
        if (want_db) {  // item m0 + t: sum of its 32 entries of the chunk, fixed order (hi + lo == g exactly)
            const uint32_t rb = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const uint32_t off = rb + ((c4 ^ (t & 7)) << 4);
                const float4 h = *reinterpret_cast<const float4*>(c.A_hi + off);
                float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
                if (passes == 3) l = *reinterpret_cast<const float4*>(c.A_lo + off);
                db_acc += (h.x + l.x);
                db_acc += (h.y + l.y);
                db_acc += (h.z + l.z);
                db_acc += (h.w + l.w);
            }
        }
        wait(c);
    }
    const bool row_ok = m0 + t < p.n_dec;
    float* out = p.dW4 + (int64_t)(m0 + t) * p.H + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
        float v[32];
        if (M > 0) {
            load_acc32(c, c0, v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) st4(out + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
    }
    if (want_db && row_ok) p.db4[m0 + t] = db_acc;
    teardown(c);
}

// ---------------------------------------------------------------- launchers
static int set_smem(const void* fn, int bytes) {
    DMT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

int decoder_tc_splits(int n_dec, int chunks_per_split) {
    const int n_chunks = (n_dec + TK - 1) / TK;
    return (n_chunks + chunks_per_split - 1) / chunks_per_split;
}
// k-chunks per split so that one organization-step launches about one wave of CTAs, at most kMaxSplits slices
int decoder_tc_chunks_per_split(int n_rows_max, int n_dec, int H) {
    const int n_chunks = (n_dec + TK - 1) / TK;
    const int tiles = ((n_rows_max + TM - 1) / TM) * (H / TN);
    int splits = (kNumSMs + tiles - 1) / tiles;
    if (splits > 16) splits = 16;
    if (splits > n_chunks) splits = n_chunks;
    if (splits < 1) splits = 1;
    return (n_chunks + splits - 1) / splits;
}
int64_t decoder_tc_scratch_floats(int n_rows_max, int n_dec, int H) {
    const int cps = decoder_tc_chunks_per_split(n_rows_max, n_dec, H);
    const int64_t tiles = (int64_t)((n_rows_max + TM - 1) / TM) * ((n_dec + TN - 1) / TN);
    return (int64_t)decoder_tc_splits(n_dec, cps) * n_rows_max * H + tiles + 64;
}

int launch_decoder_tc_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                          const float* A3, const float* W4, const float* b4, int H, int n_dec, int loss_kind,
                          const int32_t* n_targets, const int32_t* ent_off, float* pred, float* gout, float* loss_part,
                          int passes, int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(reinterpret_cast<const void*>(dec_fwd_tc_kernel), smem_bytes(kFwdExtra));
        if (rc) return rc;
        configured = true;
    }
    DecTcArgs a{rows, indptr, indices, target, A3, W4, b4, H, n_dec, loss_kind, n_targets, ent_off, gout, pred, loss_part};
    dim3 grid((n_dec + TN - 1) / TN, (n_rows_max + TM - 1) / TM);
    dec_fwd_tc_kernel<<<grid, kThreads, smem_bytes(kFwdExtra), st>>>(a, br, passes);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_decoder_tc_bwd_a(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* gbuf,
                            const int32_t* ent_off, const float* A3, const float* W4, int H, int n_dec, float* part,
                            const float* loss_part, float* dZ3, float* loss_rows, int tanh_deriv, int passes,
                            int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(reinterpret_cast<const void*>(dec_bwd_a_tc_kernel), smem_bytes(0));
        if (rc) return rc;
        configured = true;
    }
    const int cps = decoder_tc_chunks_per_split(n_rows_max, n_dec, H);
    const int splits = decoder_tc_splits(n_dec, cps);
    DecBwdAArgs a{rows, indptr, indices, gbuf, ent_off, W4, H, n_dec, cps, part, n_rows_max};
    dim3 grid(H / TN, (n_rows_max + TM - 1) / TM, splits);
    dec_bwd_a_tc_kernel<<<grid, kThreads, smem_bytes(0), st>>>(a, br, passes);
    DMT_LAUNCH_CHECK();
    dec_bwd_a_finish_kernel<<<n_rows_max, 256, 0, st>>>(part, splits, n_rows_max, A3, H, tanh_deriv, dZ3, loss_part,
                                                        (n_dec + TN - 1) / TN, loss_rows, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_decoder_tc_bwd_w(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* gbuf,
                            const int32_t* ent_off, const float* A3, int H, int n_dec, float* dW4, float* db4,
                            int passes, int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    const int extra = align16(3 * n_rows_max * 4);
    if (smem_bytes(extra) > 227 * 1024) {
        set_error("decoder tensor-core path: batch too large for the per-row index arrays in shared memory");
        return DMT_E_ARG;
    }
    static int configured = 0;
    if (configured < smem_bytes(extra)) {
        int rc = set_smem(reinterpret_cast<const void*>(dec_bwd_w_tc_kernel), smem_bytes(extra));
        if (rc) return rc;
        configured = smem_bytes(extra);
    }
    DecBwdWArgs a{rows, indptr, indices, gbuf, ent_off, A3, H, n_dec, dW4, db4, n_rows_max};
    dim3 grid(H / TN, (n_dec + TM - 1) / TM);
    dec_bwd_w_tc_kernel<<<grid, kThreads, smem_bytes(extra), st>>>(a, br, passes);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int64_t dmt_ae_decoder_tc_scratch_floats(int n_rows, int n_dec, int H) {
    if (n_rows <= 0 || n_dec <= 0 || H <= 0 || H % 128 != 0) return 0;
    return decoder_tc_scratch_floats(n_rows, n_dec, H);
}

int dmt_ae_decoder_tc(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices,
                      const float* target, const float* A3, const float* W4, const float* b4, int H, int n_dec,
                      int loss_kind, const int32_t* n_targets, int passes, float* pred, float* gout, float* dZ3,
                      float* dW4, float* db4, float* loss_rows, int tanh_deriv, float* scratch, void* stream) {
    DMT_REQUIRE(n_rows >= 0 && n_dec > 0 && H > 0 && H % 128 == 0, "dmt_ae_decoder_tc: H must be a multiple of 128");
    DMT_REQUIRE(passes == 1 || passes == 3, "dmt_ae_decoder_tc: passes must be 1 or 3");
    const bool train = gout != nullptr;
    DMT_REQUIRE(!train || (n_targets && target && dZ3 && dW4 && loss_rows && scratch),
                "dmt_ae_decoder_tc: train mode needs n_targets, target, dZ3, dW4, loss_rows and scratch");
    if (n_rows == 0) return 0;
    cudaStream_t st = as_stream(stream);
    BatchRef br = batch_by_value(0, n_rows);
    const int cps = decoder_tc_chunks_per_split(n_rows, n_dec, H);
    float* part = scratch;
    float* loss_part = scratch ? scratch + (int64_t)decoder_tc_splits(n_dec, cps) * n_rows * H : nullptr;
    int rc = launch_decoder_tc_fwd(rows, indptr, indices, target, A3, W4, b4, H, n_dec, loss_kind, n_targets, nullptr,
                                   pred, gout, loss_part, passes, n_rows, br, st);
    if (rc || !train) return rc;
    if ((rc = launch_decoder_tc_bwd_a(rows, indptr, indices, gout, nullptr, A3, W4, H, n_dec, part, loss_part, dZ3,
                                      loss_rows, tanh_deriv, passes, n_rows, br, st)))
        return rc;
    return launch_decoder_tc_bwd_w(rows, indptr, indices, gout, nullptr, A3, H, n_dec, dW4, db4, passes, n_rows, br, st);
}

}  // extern "C"
