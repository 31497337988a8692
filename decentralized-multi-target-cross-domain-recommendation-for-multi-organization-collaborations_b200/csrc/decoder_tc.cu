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
// D1's own area behind the ring: row offsets [TM+1], row starts [TM], output bases [TM], reduction scratch [32]
constexpr int kFwdExtra = align16((TM + 1) * 4 + 2 * TM * 4 + 32 * 4);
static_assert(TM * kSStride * 4 <= kRingBytes, "accumulator tile must fit the ring it aliases");

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
    TileTab tab;
};

// ---------------------------------------------------------------- D1: forward + masked loss epilogue
__global__ void __launch_bounds__(kThreads, 1) dec_fwd_tc_kernel(DecTcArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    if (m0 >= M) return;
    const Pipe pp = pipe_setup(smem_raw, kFwdExtra, p.H / TK);
    const int warp = threadIdx.x >> 5;
    int32_t* s_off = reinterpret_cast<int32_t*>(pp.extra);  // [TM + 1] counts, then their exclusive scan
    int32_t* s_start = s_off + TM + 1;                       // [TM]
    int32_t* s_obase = s_start + TM;                         // [TM]
    float* s_red = reinterpret_cast<float*>(s_obase + TM);   // [32]
    float* S = reinterpret_cast<float*>(pp.base);            // accumulator tile, aliases the ring after the main loop
    if (warp < 8) {
        const int team = warp >> 2, tt = threadIdx.x & (kTeam - 1);
        const int n_mine = team == 0 ? pp.n0 : pp.n1, first = team == 0 ? 0 : pp.n0;
        for (int i = 0; i < n_mine; ++i) {
            const int c = chunk_slot(pp, team, i), k0 = (first + i) * TK;
            const Stage st = producer_acquire(pp, c);
            stage_kcontig(tt, p.A3, p.H, m0, M, k0, p.H, st.A_hi, st.A_lo, passes);
            stage_kcontig(tt, p.W4, p.H, n0, p.n_dec, k0, p.H, st.B_hi, st.B_lo, passes);
            producer_commit(pp, c);
        }
        wait_accumulator(pp);
        const int row = epi_row(), c0 = epi_col0();
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            load_acc32(pp, c0 + cc, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) S[row * kSStride + c0 + cc + j] = v[j];
        }
    } else if (warp == kMmaWarp) {
        mma_loop(pp, passes);
    } else {
        // auxiliary warp: this tile's window of target entries in every row, while the main loop runs
        for (int r = threadIdx.x & 31; r < TM; r += 32) {
            int e0 = 0, s = 0, e = 0, obase = 0;
            if (m0 + r < M) {
                const int u = p.rows[lo + m0 + r];
                row_window(p.tab, p.indptr, p.indices, lo + m0 + r, u, blockIdx.x, blockIdx.x + 1, e0, s, e);
                obase = p.ent_off ? p.ent_off[lo + m0 + r] - e0 : 0;
            }
            s_off[r] = e - s;
            s_start[r] = s;
            s_obase[r] = obase;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the 128 counts: 4 per lane
        const int t = threadIdx.x;
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
    for (int j = threadIdx.x; j < total; j += kThreads) {  // the tile's target entries, flattened over all threads
        int rl = 0, rh = TM;                                // row r with s_off[r] <= j < s_off[r + 1]
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
        if (threadIdx.x == 0) p.loss_part[blockIdx.y * gridDim.x + blockIdx.x] = loss_acc;
    }
    pipe_teardown(pp);
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
    int chunks_per_split;    // k-chunks (of 32 items) per grid.z slice; a multiple of 8 (team halves on tile borders)
    float* part;             // [grid.z][part_rows x H]
    int part_rows;
    TileTab tab;
};

__global__ void __launch_bounds__(kThreads, 1) dec_bwd_a_tc_kernel(DecBwdAArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    if (m0 >= M) return;
    const int n_chunks_all = (p.n_dec + TK - 1) / TK;
    const int kc0 = blockIdx.z * p.chunks_per_split;
    const int kc1 = min(n_chunks_all, kc0 + p.chunks_per_split);
    // team 0 takes the first chunks_per_split/2 chunks of the slice (a multiple of 4: 128-column tile borders)
    const int half = p.chunks_per_split >> 1;
    Pipe pp = pipe_setup(smem_raw, 0, kc1 - kc0);
    pp.n0 = min(half, kc1 - kc0);
    pp.n1 = (kc1 - kc0) - pp.n0;
    const int warp = threadIdx.x >> 5;
    if (warp < 8) {
        const int team = warp >> 2, tt = threadIdx.x & (kTeam - 1);
        const int n_mine = team == 0 ? pp.n0 : pp.n1, first = kc0 + (team == 0 ? 0 : pp.n0);
        int ptr = 0, e1 = 0;
        const float* grow = p.gbuf;
        if (m0 + tt < M && n_mine > 0) {  // this thread walks row m0 + tt of G through its team's item range
            const int u = p.rows[lo + m0 + tt];
            int e0, e;
            row_window(p.tab, p.indptr, p.indices, lo + m0 + tt, u, first / 4, (first + n_mine + 3) / 4, e0, ptr, e);
            e1 = p.indptr[u + 1];
            if (p.ent_off) grow = p.gbuf + ((int64_t)p.ent_off[lo + m0 + tt] - e0);
        }
        for (int i = 0; i < n_mine; ++i) {
            const int c = chunk_slot(pp, team, i), k0 = (first + i) * TK;
            const Stage st = producer_acquire(pp, c);
            zero_tiles(tt, st.A_hi, st.A_lo, passes);
            // B(n = unit, k = item) = W4[item][unit]
            stage_transposed(tt, p.W4, p.H, n0, p.H, k0, p.n_dec, st.B_hi, st.B_lo, passes);
            team_sync(team);  // canvas zeroed by the whole team before any scatter
            while (ptr < e1) {
                const int col = p.indices[ptr];
                if (col >= k0 + TK) break;
                store_split1(st.A_hi, st.A_lo, tile_off(tt, col - k0), grow[ptr], passes);
                ++ptr;
            }
            producer_commit(pp, c);
        }
        wait_accumulator(pp);
        const int row = epi_row(), c0 = epi_col0();
        float* out = p.part + ((int64_t)blockIdx.z * p.part_rows + (m0 + row)) * p.H + n0 + c0;
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            load_acc32(pp, c0 + cc, v);  // warp-collective: rows past the batch just do not store
            if (m0 + row < M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) st4(out + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            }
        }
    } else if (warp == kMmaWarp) {
        mma_loop(pp, passes);
    }
    pipe_teardown(pp);
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
    TileTab tab;
};

__global__ void __launch_bounds__(kThreads, 1) dec_bwd_w_tc_kernel(DecBwdWArgs p, BatchRef br, int passes) {
    extern __shared__ uint8_t smem_raw[];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    const int M = hi - lo;
    const int m0 = blockIdx.y * TM /* item tile */, n0 = blockIdx.x * TN /* hidden units */;
    const int extra = align16(3 * p.rows_cap * 4);
    const Pipe pp = pipe_setup(smem_raw, extra, (M + TK - 1) / TK);
    const int warp = threadIdx.x >> 5;
    int32_t* r_start = reinterpret_cast<int32_t*>(pp.extra);  // first entry of the row inside the item tile
    int32_t* r_end = r_start + p.rows_cap;
    int32_t* r_gofs = r_end + p.rows_cap;  // gbuf index of entry e is r_gofs + e
    if (warp < 8) {
        const int team = warp >> 2, tt = threadIdx.x & (kTeam - 1);
        for (int r = threadIdx.x; r < M; r += kProducers) {
            const int u = p.rows[lo + r];
            int e0, s, e;
            row_window(p.tab, p.indptr, p.indices, lo + r, u, blockIdx.y, blockIdx.y + 1, e0, s, e);
            r_start[r] = s;
            r_end[r] = e;
            r_gofs[r] = p.ent_off ? p.ent_off[lo + r] - e0 : 0;
        }
        producers_sync();
        const bool want_db = p.db4 != nullptr && blockIdx.x == 0;
        float db_acc = 0.f;
        const int n_mine = team == 0 ? pp.n0 : pp.n1, first = team == 0 ? 0 : pp.n0;
        for (int i = 0; i < n_mine; ++i) {
            const int c = chunk_slot(pp, team, i), k0 = (first + i) * TK;
            const Stage st = producer_acquire(pp, c);
            zero_tiles(tt, st.A_hi, st.A_lo, passes);
            // B(n = unit, k = batch row) = A3[row][unit]
            stage_transposed(tt, p.A3, p.H, n0, p.H, k0, M, st.B_hi, st.B_lo, passes);
            team_sync(team);
            {   // four threads per batch row of the chunk scatter its in-tile entries: tile(item - m0, row - k0) = g
                const int r = k0 + (tt >> 2);
                if (r < M) {
                    const int e_end = r_end[r], gofs = r_gofs[r];
                    for (int e = r_start[r] + (tt & 3); e < e_end; e += 4)
                        store_split1(st.A_hi, st.A_lo, tile_off(p.indices[e] - m0, r - k0), p.gbuf[(int64_t)gofs + e],
                                     passes);
                }
            }
            if (want_db) {  // item m0 + tt: sum of its 32 entries of the chunk in row order (hi + lo == g exactly)
                team_sync(team);
                const uint32_t rb = (uint32_t)((tt >> 3) * 1024 + (tt & 7) * 128);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const uint32_t off = rb + ((c4 ^ (tt & 7)) << 4);
                    const float4 h = *reinterpret_cast<const float4*>(st.A_hi + off);
                    float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (passes == 3) l = *reinterpret_cast<const float4*>(st.A_lo + off);
                    db_acc += (h.x + l.x);
                    db_acc += (h.y + l.y);
                    db_acc += (h.z + l.z);
                    db_acc += (h.w + l.w);
                }
            }
            producer_commit(pp, c);
        }
        if (want_db) {  // the two teams summed disjoint halves of the batch rows: team 0's half first
            float* s_db = reinterpret_cast<float*>(pp.extra);  // reuses r_start (all scatters are done)
            producers_sync();
            if (team == 1) s_db[tt] = db_acc;
            producers_sync();
            if (team == 0 && m0 + tt < p.n_dec) p.db4[m0 + tt] = db_acc + s_db[tt];
        }
        wait_accumulator(pp);
        const int row = epi_row(), c0 = epi_col0();
        const bool row_ok = m0 + row < p.n_dec;
        float* out = p.dW4 + (int64_t)(m0 + row) * p.H + n0 + c0;
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            if (M > 0) {
                load_acc32(pp, c0 + cc, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) st4(out + cc + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            }
        }
    } else if (warp == kMmaWarp) {
        mma_loop(pp, passes);
    }
    pipe_teardown(pp);
}

// tab[r * (n_tiles + 1) + j] = first entry of CSR row rows[r] (rows == null: row r) with column >= 128 j
__global__ void __launch_bounds__(256) tile_tab_kernel(const int32_t* __restrict__ rows, int n_rows,
                                                       const int32_t* __restrict__ indptr,
                                                       const int32_t* __restrict__ indices, int n_tiles,
                                                       int32_t* __restrict__ tab) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rows * (n_tiles + 1)) return;
    const int r = (int)(i / (n_tiles + 1)), j = (int)(i % (n_tiles + 1));
    const int u = rows ? rows[r] : r;
    const int e0 = indptr[u], e1 = indptr[u + 1];
    tab[i] = j == n_tiles ? e1 : lower_bound_i32(indices, e0, e1, j * TN);
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
    // a multiple of 8: each producer team takes half a slice, and halves start on 128-column tile borders
    return ((n_chunks + splits - 1) / splits + 7) / 8 * 8;
}
int64_t decoder_tc_scratch_floats(int n_rows_max, int n_dec, int H) {
    const int cps = decoder_tc_chunks_per_split(n_rows_max, n_dec, H);
    const int64_t tiles = (int64_t)((n_rows_max + TM - 1) / TM) * ((n_dec + TN - 1) / TN);
    return (int64_t)decoder_tc_splits(n_dec, cps) * n_rows_max * H + tiles + 64;
}
int64_t decoder_tc_tab_ints(int64_t n_rows, int n_dec) { return n_rows * ((n_dec + TN - 1) / TN + 1); }

int build_tile_tab(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices, int n_dec,
                   int32_t* tab, cudaStream_t st) {
    const int n_tiles = (n_dec + TN - 1) / TN;
    const int64_t n = (int64_t)n_rows * (n_tiles + 1);
    if (n <= 0) return 0;
    tile_tab_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows, n_rows, indptr, indices, n_tiles, tab);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_decoder_tc_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                          const float* A3, const float* W4, const float* b4, int H, int n_dec, int loss_kind,
                          const int32_t* n_targets, const int32_t* ent_off, float* pred, float* gout, float* loss_part,
                          TcTab tab, int passes, int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    static bool configured = false;
    if (!configured) {
        int rc = set_smem(reinterpret_cast<const void*>(dec_fwd_tc_kernel), smem_bytes(kFwdExtra));
        if (rc) return rc;
        configured = true;
    }
    DecTcArgs a{rows, indptr, indices, target, A3, W4, b4, H, n_dec, loss_kind, n_targets, ent_off, gout, pred, loss_part,
                TileTab{tab.tab, tab.tab ? tab.mode : 0, (n_dec + TN - 1) / TN}};
    dim3 grid((n_dec + TN - 1) / TN, (n_rows_max + TM - 1) / TM);
    dec_fwd_tc_kernel<<<grid, kThreads, smem_bytes(kFwdExtra), st>>>(a, br, passes);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_decoder_tc_bwd_a(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* gbuf,
                            const int32_t* ent_off, const float* A3, const float* W4, int H, int n_dec, float* part,
                            const float* loss_part, float* dZ3, float* loss_rows, int tanh_deriv, TcTab tab, int passes,
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
    DecBwdAArgs a{rows, indptr, indices, gbuf, ent_off, W4, H, n_dec, cps, part, n_rows_max,
                  TileTab{tab.tab, tab.tab ? tab.mode : 0, (n_dec + TN - 1) / TN}};
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
                            TcTab tab, int passes, int n_rows_max, BatchRef br, cudaStream_t st) {
    if (n_rows_max <= 0) return 0;
    const int extra = align16(3 * n_rows_max * 4);
    if (smem_bytes(extra) > 227 * 1024) {  // ~2800 rows
        set_error("decoder tensor-core path: batch too large for the per-row index arrays in shared memory");
        return DMT_E_ARG;
    }
    static int configured = 0;
    if (configured < smem_bytes(extra)) {
        int rc = set_smem(reinterpret_cast<const void*>(dec_bwd_w_tc_kernel), smem_bytes(extra));
        if (rc) return rc;
        configured = smem_bytes(extra);
    }
    DecBwdWArgs a{rows, indptr, indices, gbuf, ent_off, A3, H, n_dec, dW4, db4, n_rows_max,
                  TileTab{tab.tab, tab.tab ? tab.mode : 0, (n_dec + TN - 1) / TN}};
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
    return decoder_tc_scratch_floats(n_rows, n_dec, H) + decoder_tc_tab_ints(n_rows, n_dec);
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
    // per-batch-row table of the CSR windows at 128-column tile borders (the engine keeps one per CSR instead)
    TcTab tab{nullptr, 0};
    int rc = 0;
    if (scratch) {
        int32_t* t = reinterpret_cast<int32_t*>(scratch + decoder_tc_scratch_floats(n_rows, n_dec, H));
        if ((rc = build_tile_tab(rows, n_rows, indptr, indices, n_dec, t, st))) return rc;
        tab = TcTab{t, 1};
    }
    rc = launch_decoder_tc_fwd(rows, indptr, indices, target, A3, W4, b4, H, n_dec, loss_kind, n_targets, nullptr, pred,
                               gout, loss_part, tab, passes, n_rows, br, st);
    if (rc || !train) return rc;
    if ((rc = launch_decoder_tc_bwd_a(rows, indptr, indices, gout, nullptr, A3, W4, H, n_dec, part, loss_part, dZ3,
                                      loss_rows, tanh_deriv, tab, passes, n_rows, br, st)))
        return rc;
    return launch_decoder_tc_bwd_w(rows, indptr, indices, gout, nullptr, A3, H, n_dec, dW4, db4, tab, passes, n_rows, br,
                                   st);
}

}  // extern "C"
