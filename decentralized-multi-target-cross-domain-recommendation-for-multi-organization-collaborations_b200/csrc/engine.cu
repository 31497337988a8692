// Device-resident organization engine: Organization.train / Organization.predict of the reference
// (src/organization.py:140-217) for the AAE, with the batch loop moved onto the device.
//
// Per local epoch (one call):
//   plan   : from the epoch's row order build, for targets and data, the batch-ordered entry space, its stable
//            sort by (batch, column) and the (batch, column) segments  -> no per-batch host work at all
//   steps  : for every batch one fixed sequence of 20 kernels (forward, fused decoder/loss/first
//            backward product, segmented gradient reductions, dense backward, global-norm clip + Adam); every
//            kernel reads its batch bounds from device memory, so the whole epoch is ONE CUDA graph that is
//            captured once and replayed for every epoch of every round.
// Organizations own private streams: the per-organization graphs of one rank run concurrently and fill the
// 148 SMs even though a single 500-row batch cannot.
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "kernels.cuh"

namespace dmt {

constexpr int kForkEvents = 8;

struct PlanSide {  // per-matrix (targets or data) epoch plan buffers
    int64_t cap = 0;                // entry capacity
    int32_t* ent_off = nullptr;     // [rows_cap + 1] offsets of each batch-row in the batch-ordered entry space
    int32_t* len = nullptr;         // [rows_cap + 1]
    uint32_t* key = nullptr;        // [cap] batch * n_cols + col
    int32_t* ent_row = nullptr;     // [cap] row index inside its batch
    int32_t* perm = nullptr;        // [cap] entry ids sorted by key (stable)
    int32_t* seg_key = nullptr;     // [cap]
    int32_t* seg_off = nullptr;     // [cap + 1]
    int32_t* n_seg = nullptr;       // [1]
    int32_t* batch_seg_off = nullptr;  // [nb_cap + 1]
    int32_t* batch_cnt = nullptr;   // [nb_cap] entries per batch
    // load-balanced segmented reduction: segments cut into chunks of <= kSegChunk entries
    int32_t* n_ch = nullptr;           // [cap + 2]
    int32_t* seg_chunk_off = nullptr;  // [cap + 2]
    int32_t* chunk_seg = nullptr;      // [chunk_cap]
    int32_t* batch_chunk_off = nullptr;  // [nb_cap + 2]
    float* part = nullptr;             // [part_rows x H1]
    float* part_bias = nullptr;        // [part_rows]
    int64_t chunk_cap = 0, part_rows = 0;
};

}  // namespace dmt

using namespace dmt;

struct dmt_org {
    int n_rows, n_enc, n_dec, H1, H2, batch_rows, loss_kind;
    cudaStream_t st;
    bool own_stream;
    const int32_t *d_indptr, *d_indices, *t_indptr, *t_indices;
    const float *d_val, *t_val;
    int64_t d_nnz, t_nnz;
    // flat parameter / gradient / moment buffers and sub-views
    int64_t n_params;
    float *P, *G, *M, *V;
    int64_t oW1, ob1, oW2, ob2, oW3, ob3, oW4, ob4;
    // activations
    int act_rows;
    float *a1, *a2, *c, *a3, *dz3, *dz2, *dz1, *loss_rows;
    int32_t* iota_rows;
    // plan
    int rows_cap, nb_cap;
    PlanSide pt, pd;
    float* gbuf;      // [t cap] dL/do per target entry (batch-ordered)
    float* dval_ord;  // [d cap] data values in batch-ordered entry space
    int32_t* row_batch;  // [rows_cap]
    int32_t* active;     // [nb_cap]
    // decoder row chunks (<= kDecChunk targets each)
    int32_t *t_nch_row, *t_chunk_off, *t_chunk_row, *t_batch_chunk;
    float *dz_part, *loss_part;
    int64_t dec_chunk_cap, dec_part_rows;
    // tensor-core decoder (decoder_tc.cu)
    int dec_mode, dec_passes;
    int dec_blocks;  // grid of the decoder chunk kernel (0: two per SM), dmt_org_set_decoder_blocks
    int gather;      // fused step: 1 = weight / activation rows through bulk-copy rings (bulk.cuh), 0 = plain loads
    int bulk_blocks; // grid of the bulk decoder kernel (0: five per SM)
    int dec_form;    // register-load decoder of the fused step: 2 = four-warp blocks + next-chunk prefetch, 0 = eight warps
    int pdl;         // fused step: same-stream kernels launched with programmatic stream serialization (common.cuh)
    int rows_mode;   // fused step, row kernels: 1 = W2 / W3 streamed through shared memory by bulk copies (fused_rows.cu)
    int rows_R;      // their row tile (fused_rows_per_cta)
    int stream_R;    // row tile of the register-streamed row kernels (4, 8 or 16)
    int fanout;  // 1: the backward pass of a step is enqueued as parallel branches (dmt_org_set_fanout)
    float* tc_scratch;  // split-K partials [splits x batch_rows x H1], then per-tile loss sums
    // tables of per-row CSR windows at 128-column tile borders, one per target CSR seen (train targets, predict splits)
    struct TabEntry { const int32_t *indptr, *indices; int n_rows; int32_t* tab; };
    TabEntry tabs[4];
    int n_tabs;
    TcTab train_tab;  // table of the training target CSR (set when the tensor-core mode is switched on)
    void* sort_temp;
    int64_t sort_temp_bytes;
    // epoch inputs (stable addresses for the graph)
    int32_t* rows_buf;     // [rows_cap]
    int32_t* row_off_buf;  // [nb_cap + 1]
    uint8_t* keep_buf;     // [rows_cap * H2]
    uint64_t* seed_dev;
    float* loss_buf;  // [nb_cap]
    // optimizer scalars
    float* partial;
    AdamScalars* sc;
    int* step_dev;
    cudaEvent_t ev;
    // backward fan-out: auxiliary streams and fork/join events (parallel branches of the captured epoch graph)
    cudaStream_t aux[3];
    cudaEvent_t fev[8];
    // fused six-launch step (fused.cu): transposed shadows of W2 / W3, plan-time metadata, arrival counters
    int step_mode;  // 1: fused (default when H1 = 256, H2 = 128 and the gather decoder), 0: classic 20-kernel step
    float *W2t, *W3t;
    int4* dec_meta;
    int32_t* row_cnt;
    float* g_sorted;
    int4 *t_seg_meta, *d_seg_meta;
    int32_t *t_row_sorted, *d_row_sorted, *t_inv_perm;
    float* d_val_sorted;
    int32_t *t_seg_cnt, *d_seg_cnt;
    float *part_db, *part_dw;
    int32_t *dw_cnt, *norm_ticket;
    // graph cache
    cudaGraphExec_t exec;
    long long g_kernels;  // our kernel launches captured in the graph
    int g_nb, g_keep, g_rows;
    int64_t g_nt, g_nd;
    AdamHyper g_hp;
};

namespace dmt {

static int free_all(dmt_org* o);

template <class T>
static int dalloc(T** p, int64_t n) {
    if (n < 1) n = 1;
    DMT_CUDA(cudaMalloc(reinterpret_cast<void**>(p), (size_t)n * sizeof(T)));
    return 0;
}

static int alloc_side(PlanSide& s, int64_t cap, int rows_cap, int nb_cap, int n_cols, int H, int batch_rows) {
    s.cap = cap;
    s.chunk_cap = cap + cap / kSegChunk + 16;
    int64_t per_batch = (int64_t)batch_rows * n_cols < cap ? (int64_t)batch_rows * n_cols : cap;  // entries of one batch
    s.part_rows = (int64_t)n_cols + per_batch / kSegChunk + 16;
    int rc = 0;
    if ((rc = dalloc(&s.n_ch, cap + 2))) return rc;
    if ((rc = dalloc(&s.seg_chunk_off, cap + 2))) return rc;
    if ((rc = dalloc(&s.chunk_seg, s.chunk_cap))) return rc;
    if ((rc = dalloc(&s.batch_chunk_off, nb_cap + 2))) return rc;
    if ((rc = dalloc(&s.part, s.part_rows * H))) return rc;
    if ((rc = dalloc(&s.part_bias, s.part_rows))) return rc;
    if ((rc = dalloc(&s.ent_off, rows_cap + 2))) return rc;
    if ((rc = dalloc(&s.len, rows_cap + 2))) return rc;
    if ((rc = dalloc(&s.key, cap))) return rc;
    if ((rc = dalloc(&s.ent_row, cap))) return rc;
    if ((rc = dalloc(&s.perm, cap))) return rc;
    if ((rc = dalloc(&s.seg_key, cap))) return rc;
    if ((rc = dalloc(&s.seg_off, cap + 2))) return rc;
    if ((rc = dalloc(&s.n_seg, 1))) return rc;
    if ((rc = dalloc(&s.batch_seg_off, nb_cap + 2))) return rc;
    if ((rc = dalloc(&s.batch_cnt, nb_cap + 1))) return rc;
    return 0;
}

static void free_side(PlanSide& s) {
    cudaFree(s.ent_off); cudaFree(s.len); cudaFree(s.key); cudaFree(s.ent_row); cudaFree(s.perm);
    cudaFree(s.seg_key); cudaFree(s.seg_off); cudaFree(s.n_seg); cudaFree(s.batch_seg_off); cudaFree(s.batch_cnt);
    cudaFree(s.n_ch); cudaFree(s.seg_chunk_off); cudaFree(s.chunk_seg); cudaFree(s.batch_chunk_off);
    cudaFree(s.part); cudaFree(s.part_bias);
}

// ---------------------------------------------------------------- plan kernels
__global__ void plan_row_len_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ row_off, int nb,
                                    int n, const int32_t* __restrict__ t_indptr,
                                    const int32_t* __restrict__ d_indptr, int32_t* __restrict__ tlen,
                                    int32_t* __restrict__ dlen, int32_t* __restrict__ row_batch) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    if (j == n) {  // terminator for the exclusive scan
        tlen[j] = 0;
        dlen[j] = 0;
        return;
    }
    int u = rows[j];
    tlen[j] = t_indptr[u + 1] - t_indptr[u];
    dlen[j] = d_indptr[u + 1] - d_indptr[u];
    int lo = 0, hi = nb;  // batch of row j: largest b with row_off[b] <= j
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (row_off[mid] <= j) lo = mid; else hi = mid;
    }
    row_batch[j] = lo;
}

__global__ void plan_batch_meta_kernel(const int32_t* __restrict__ row_off, int nb,
                                       const int32_t* __restrict__ t_off, const int32_t* __restrict__ d_off,
                                       int32_t* __restrict__ t_cnt, int32_t* __restrict__ d_cnt,
                                       int32_t* __restrict__ active) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int r0 = row_off[b], r1 = row_off[b + 1];
    t_cnt[b] = t_off[r1] - t_off[r0];
    d_cnt[b] = d_off[r1] - d_off[r0];
    // a batch without DATA entries is skipped entirely (reference src/organization.py:153-155)
    active[b] = d_cnt[b] > 0 ? 1 : 0;
}

// one warp per batch-row: writes the (batch, column) keys and the in-batch row index of every entry
__global__ void __launch_bounds__(256) plan_fill_kernel(const int32_t* __restrict__ rows,
                                                        const int32_t* __restrict__ row_off,
                                                        const int32_t* __restrict__ row_batch, int n,
                                                        const int32_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ indices,
                                                        const float* __restrict__ val, int n_cols,
                                                        const int32_t* __restrict__ ent_off,
                                                        uint32_t* __restrict__ key, int32_t* __restrict__ ent_row,
                                                        float* __restrict__ val_ord) {
    int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (j >= n) return;
    int u = rows[j], b = row_batch[j];
    int s = indptr[u], len = indptr[u + 1] - s, o = ent_off[j];
    int rin = j - row_off[b];
    for (int k = lane; k < len; k += 32) {
        key[o + k] = (uint32_t)b * (uint32_t)n_cols + (uint32_t)indices[s + k];
        ent_row[o + k] = rin;
        if (val_ord != nullptr) val_ord[o + k] = val[s + k];
    }
}

__global__ void plan_batch_seg_kernel(const int32_t* __restrict__ seg_key, const int32_t* __restrict__ n_seg, int nb,
                                      int n_cols, int32_t* __restrict__ batch_seg_off) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    int ns = n_seg[0];
    uint32_t bound = (uint32_t)b * (uint32_t)n_cols;
    int lo = 0, hi = ns;  // first segment with key >= bound
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((uint32_t)seg_key[mid] < bound) lo = mid + 1; else hi = mid;
    }
    batch_seg_off[b] = lo;
}

__global__ void plan_row_chunks_kernel(const int32_t* __restrict__ tlen, int n, int32_t* __restrict__ nch) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    nch[j] = j < n ? (tlen[j] + kDecChunk - 1) / kDecChunk : 0;
}

__global__ void plan_row_chunk_fill_kernel(const int32_t* __restrict__ chunk_off, int n,
                                           int32_t* __restrict__ chunk_row) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int c = chunk_off[j]; c < chunk_off[j + 1]; ++c) chunk_row[c] = j;
}

// out[b] = src[idx[b]] for b in [0, nb]
__global__ void plan_gather_offsets_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int nb,
                                           int32_t* __restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= nb) out[b] = src[idx[b]];
}

static int bits_for(int64_t bound) {
    int bits = 1;
    while (bits < 32 && (1LL << bits) < bound) ++bits;
    return bits;
}

static int scan_lengths(dmt_org* o, PlanSide& s, int n, cudaStream_t st) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, s.len, s.ent_off, n + 1);
    if ((int64_t)bytes > o->sort_temp_bytes) {
        set_error("plan: scan temp too small");
        return DMT_E_STATE;
    }
    DMT_CUDA(cub::DeviceScan::ExclusiveSum(o->sort_temp, bytes, s.len, s.ent_off, n + 1, st));
    return 0;
}

static int build_plan(dmt_org* o, int n, int nb, int64_t n_t, int64_t n_d, cudaStream_t st) {
    plan_row_len_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(o->rows_buf, o->row_off_buf, nb, n, o->t_indptr,
                                                            o->d_indptr, o->pt.len, o->pd.len, o->row_batch);
    DMT_LAUNCH_CHECK();
    int rc;
    if ((rc = scan_lengths(o, o->pt, n, st))) return rc;
    if ((rc = scan_lengths(o, o->pd, n, st))) return rc;
    plan_batch_meta_kernel<<<(nb + 127) / 128, 128, 0, st>>>(o->row_off_buf, nb, o->pt.ent_off, o->pd.ent_off,
                                                            o->pt.batch_cnt, o->pd.batch_cnt, o->active);
    DMT_LAUNCH_CHECK();
    {   // decoder row chunks
        plan_row_chunks_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(o->pt.len, n, o->t_nch_row);
        DMT_LAUNCH_CHECK();
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, o->t_nch_row, o->t_chunk_off, n + 1);
        DMT_CUDA(cub::DeviceScan::ExclusiveSum(o->sort_temp, bytes, o->t_nch_row, o->t_chunk_off, n + 1, st));
        if (n > 0) {
            plan_row_chunk_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(o->t_chunk_off, n, o->t_chunk_row);
            DMT_LAUNCH_CHECK();
        }
        plan_gather_offsets_kernel<<<(nb + 1 + 127) / 128, 128, 0, st>>>(o->t_chunk_off, o->row_off_buf, nb,
                                                                        o->t_batch_chunk);
        DMT_LAUNCH_CHECK();
    }
    int blocks = (n + 7) / 8;
    if (blocks > 0) {
        plan_fill_kernel<<<blocks, 256, 0, st>>>(o->rows_buf, o->row_off_buf, o->row_batch, n, o->t_indptr,
                                                o->t_indices, nullptr, o->n_dec, o->pt.ent_off, o->pt.key,
                                                o->pt.ent_row, nullptr);
        DMT_LAUNCH_CHECK();
        plan_fill_kernel<<<blocks, 256, 0, st>>>(o->rows_buf, o->row_off_buf, o->row_batch, n, o->d_indptr,
                                                o->d_indices, o->d_val, o->n_enc, o->pd.ent_off, o->pd.key,
                                                o->pd.ent_row, o->dval_ord);
        DMT_LAUNCH_CHECK();
    }
    if ((rc = sort_segments(o->pt.key, n_t, bits_for((int64_t)nb * o->n_dec), o->pt.perm, o->pt.seg_key,
                            o->pt.seg_off, o->pt.n_seg, o->sort_temp, o->sort_temp_bytes, st)))
        return rc;
    if ((rc = sort_segments(o->pd.key, n_d, bits_for((int64_t)nb * o->n_enc), o->pd.perm, o->pd.seg_key,
                            o->pd.seg_off, o->pd.n_seg, o->sort_temp, o->sort_temp_bytes, st)))
        return rc;
    plan_batch_seg_kernel<<<(nb + 1 + 127) / 128, 128, 0, st>>>(o->pt.seg_key, o->pt.n_seg, nb, o->n_dec,
                                                               o->pt.batch_seg_off);
    DMT_LAUNCH_CHECK();
    plan_batch_seg_kernel<<<(nb + 1 + 127) / 128, 128, 0, st>>>(o->pd.seg_key, o->pd.n_seg, nb, o->n_enc,
                                                               o->pd.batch_seg_off);
    DMT_LAUNCH_CHECK();
    PlanSide* sides[2] = {&o->pt, &o->pd};
    int64_t counts[2] = {n_t, n_d};
    for (int i = 0; i < 2; ++i) {
        PlanSide& s = *sides[i];
        if ((rc = build_seg_chunks(s.seg_off, s.n_seg, counts[i], s.n_ch, s.seg_chunk_off, s.chunk_seg, o->sort_temp,
                                   o->sort_temp_bytes, st)))
            return rc;
        plan_gather_offsets_kernel<<<(nb + 1 + 127) / 128, 128, 0, st>>>(s.seg_chunk_off, s.batch_seg_off, nb,
                                                                        s.batch_chunk_off);
        DMT_LAUNCH_CHECK();
    }
    if (o->step_mode == 1) {  // per-chunk metadata and sorted-order copies for the fused step (fused.cu)
        FusedPlanArgs a{};
        a.dec_chunk_cap = n_t / kDecChunk + n + 1;
        if (a.dec_chunk_cap > o->dec_chunk_cap) a.dec_chunk_cap = o->dec_chunk_cap;
        a.n_rows = n;
        a.t_chunk_off = o->t_chunk_off; a.t_chunk_row = o->t_chunk_row; a.rows = o->rows_buf; a.row_off = o->row_off_buf;
        a.row_batch = o->row_batch; a.t_indptr = o->t_indptr; a.t_ent_off = o->pt.ent_off; a.dec_meta = o->dec_meta;
        a.t = FusedPlanSide{n_t, o->n_dec, o->pt.n_seg, o->pt.seg_off, o->pt.seg_key, o->pt.seg_chunk_off, o->pt.perm,
                            o->pt.ent_row, nullptr, o->t_seg_meta, o->t_row_sorted, nullptr, o->t_inv_perm};
        a.d = FusedPlanSide{n_d, o->n_enc, o->pd.n_seg, o->pd.seg_off, o->pd.seg_key, o->pd.seg_chunk_off, o->pd.perm,
                            o->pd.ent_row, o->dval_ord, o->d_seg_meta, o->d_row_sorted, o->d_val_sorted, nullptr};
        if ((rc = launch_plan_fused(a, st))) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------- one training step (enqueue only)
enum StepClass { K_ENC = 0, K_DENSE_FWD, K_ZERO, K_DEC, K_SEG_W4, K_DENSE_BWD, K_SEG_W1, K_NORM, K_ADAM, K_NCLASS };

// The fused step (fused.cu): forward rows -> decoder chunks -> {backward rows | dW4 segments} ->
// {dW3 / dW2 tiles | bias gradients | dW1 segments} -> norm + scalars -> Adam. `only` profiles one launch class.
static int enqueue_step_fused(dmt_org* o, int b, bool use_keep, AdamHyper hp, int only) {
    cudaStream_t st = o->st;
    const int B = o->batch_rows;
    const BatchRef br{o->row_off_buf, o->active, b, 0, 0};
    float *W1t = o->P + o->oW1, *b1 = o->P + o->ob1, *W2 = o->P + o->oW2, *b2 = o->P + o->ob2;
    float *W3 = o->P + o->oW3, *b3 = o->P + o->ob3, *W4 = o->P + o->oW4, *b4 = o->P + o->ob4;
    float* G = o->G;
    int rc;
    Dropout drop;
    drop.keep = use_keep ? o->keep_buf : nullptr;
    drop.seed_dev = o->seed_dev;
    drop.step_dev = o->step_dev;
    drop.row_base = o->row_off_buf;
    drop.b = b;
    drop.scale = 2.0f;  // nn.Dropout(p=0.5), reference src/models/ae.py:81
    drop.p = 0.5f;
    drop.enabled = 1;
#define WANT(cls) (only < 0 || only == (cls))
    const bool pdl = only < 0 && o->pdl != 0;  // the per-class profiler measures plain launches
    if (WANT(K_ENC)) {
        FusedFwd f{br, o->rows_buf, o->d_indptr, o->d_indices, o->d_val, W1t, b1, o->W2t, b2, o->W3t, b3,
                   o->a1, o->a2, o->c, o->a3, drop};
        if ((rc = o->rows_mode ? launch_fused_fwd_tma(f, B, o->rows_R, st, pdl)
                               : launch_fused_fwd(f, B, st, pdl, o->stream_R)))
            return rc;
    }
    if (WANT(K_DEC)) {
        FusedDec d{br, o->dec_meta, o->t_batch_chunk, o->pt.batch_cnt, o->t_indices, o->t_val, o->t_inv_perm,
                   o->a3, W4, b4, o->g_sorted, o->dz3, o->loss_rows, o->dz_part, o->loss_part, o->row_cnt};
        if ((rc = launch_fused_dec(d, o->gather ? o->bulk_blocks : o->dec_blocks, o->gather ? 1 : o->dec_form, st,
                                   pdl)))
            return rc;
    }
    // dW4 / db4 only has to be complete before the norm: it runs as a parallel branch of the captured step graph next
    // to the critical path backward rows -> dW3 / dW2 / dW1 (profiling keeps everything on the main stream)
    const bool par = only < 0;
    cudaStream_t sA = par ? o->aux[0] : st;
    if (WANT(K_SEG_W4)) {
        if (par) {
            DMT_CUDA(cudaEventRecord(o->fev[0], st));
            DMT_CUDA(cudaStreamWaitEvent(sA, o->fev[0], 0));
        }
        FusedSeg s{o->t_seg_meta, o->pt.batch_chunk_off, o->t_row_sorted, o->g_sorted, o->pt.part, o->pt.part_bias,
                   o->t_seg_cnt, o->active, b};
        if ((rc = launch_fused_seg_chunks(s, o->a3, G + o->oW4, G + o->ob4, o->n_dec * 2, o->gather, sA))) return rc;
    }
    if (WANT(K_SEG_W1)) {
        FusedBwd w{br, o->pt.len, o->dz3, W3, W2, o->a1, o->a2, o->dz2, o->dz1, o->part_db, drop};
        if ((rc = o->rows_mode ? launch_fused_bwd_rows_tma(w, B, o->rows_R, st, pdl)
                               : launch_fused_bwd_rows(w, B, st, pdl, o->stream_R)))
            return rc;
    }
    if (WANT(K_DENSE_BWD)) {
        FusedGrad g{br, o->pt.len, o->dz3, o->dz2, o->c, o->a1, o->part_db, G, o->oW2, o->oW3, o->ob1, o->ob2, o->ob3,
                    o->part_dw, o->dw_cnt, o->rows_mode ? o->rows_R : o->stream_R};
        FusedSeg s{o->d_seg_meta, o->pd.batch_chunk_off, o->d_row_sorted, o->d_val_sorted, o->pd.part, o->pd.part_bias,
                   o->d_seg_cnt, o->active, b};
        if ((rc = launch_fused_grad_phase(g, s, o->dz1, G + o->oW1, o->n_enc * 2, st, pdl))) return rc;
    }
    if (par) {  // join
        DMT_CUDA(cudaEventRecord(o->fev[1], sA));
        DMT_CUDA(cudaStreamWaitEvent(st, o->fev[1], 0));
    }
    if (WANT(K_NORM))
        if ((rc = launch_norm_prepare(G, o->n_params, o->partial, o->sc, o->step_dev, o->loss_rows, o->pt.len,
                                      o->pt.batch_cnt + b, o->loss_buf + b, br, st, pdl)))
            return rc;
    if (WANT(K_ADAM))
        if ((rc = launch_adam_shadow(o->P, G, o->M, o->V, o->n_params, o->sc, hp, o->partial, o->step_dev, o->oW2,
                                     o->oW3, o->W2t, o->W3t, st, pdl)))
            return rc;
#undef WANT
    return 0;
}

static int enqueue_step(dmt_org* o, int b, bool use_keep, AdamHyper hp, int only = -1) {
    if (o->step_mode == 1 && o->dec_mode == 0) return enqueue_step_fused(o, b, use_keep, hp, only);
    cudaStream_t st = o->st;
    const int B = o->batch_rows, H1 = o->H1, H2 = o->H2;
    BatchRef br{o->row_off_buf, o->active, b, 0, 0};
    float *W1t = o->P + o->oW1, *b1 = o->P + o->ob1, *W2 = o->P + o->oW2, *b2 = o->P + o->ob2;
    float *W3 = o->P + o->oW3, *b3 = o->P + o->ob3, *W4 = o->P + o->oW4, *b4 = o->P + o->ob4;
    float* G = o->G;
    int rc;
#define WANT(cls) (only < 0 || only == (cls))
    // keep bytes of batch b start at row row_off[b]; the dense epilogue adds that base itself (replayable graph).
    Dropout drop;
    drop.keep = use_keep ? o->keep_buf : nullptr;
    drop.seed_dev = o->seed_dev;
    drop.step_dev = o->step_dev;
    drop.row_base = o->row_off_buf;
    drop.b = b;
    drop.scale = 2.0f;  // nn.Dropout(p=0.5), reference src/models/ae.py:81
    drop.p = 0.5f;
    drop.enabled = 1;
    Dropout nodrop;
    // forward
    if (WANT(K_ENC))
        if ((rc = launch_ae_encoder_fwd(o->rows_buf, o->d_indptr, o->d_indices, o->d_val, W1t, b1, H1, o->a1, B, br,
                                        st)))
            return rc;
    if (WANT(K_DENSE_FWD)) {
        if ((rc = launch_dense_fwd(o->a1, W2, b2, o->c, o->a2, drop, B, H2, H1, 1, br, st))) return rc;
        if ((rc = launch_dense_fwd(o->c, W3, b3, o->a3, nullptr, nodrop, B, H1, H2, 1, br, st))) return rc;
    }
    // (no gradient zeroing pass: Adam clears every gradient it consumes, dmt_org_set_params clears the first)
    // decoder + loss + dZ3
    const bool tc = o->dec_mode == 1;
    float* tc_loss_part = tc ? o->tc_scratch + (int64_t)decoder_tc_splits(
                                   o->n_dec, decoder_tc_chunks_per_split(B, o->n_dec, H1)) * B * H1
                             : nullptr;
    const TcTab tab = tc ? o->train_tab : TcTab{nullptr, 0};
    if (WANT(K_DEC) && tc) {
        if ((rc = launch_decoder_tc_fwd(o->rows_buf, o->t_indptr, o->t_indices, o->t_val, o->a3, W4, b4, H1, o->n_dec,
                                        DMT_LOSS_MSE, o->pt.batch_cnt, o->pt.ent_off, nullptr, o->gbuf, tc_loss_part,
                                        tab, o->dec_passes, B, br, st)))
            return rc;
        if ((rc = launch_decoder_tc_bwd_a(o->rows_buf, o->t_indptr, o->t_indices, o->gbuf, o->pt.ent_off, o->a3, W4,
                                          H1, o->n_dec, o->tc_scratch, tc_loss_part, o->dz3, o->loss_rows, 1, tab,
                                          o->dec_passes, B, br, st)))
            return rc;
    }
    if (WANT(K_DEC) && !tc) {
        DecChunks dc{o->t_chunk_off, o->t_chunk_row, o->t_batch_chunk, o->dz_part, o->loss_part};
        if ((rc = launch_ae_decoder_chunks(o->rows_buf, o->t_indptr, o->t_indices, o->t_val, o->a3, W4, b4, H1,
                                           DMT_LOSS_MSE, o->pt.batch_cnt, o->pt.ent_off, dc, o->gbuf, o->dz3,
                                           o->loss_rows, B, br, st, o->dec_blocks)))
            return rc;
    }
    // Backward fan-out. Once the decoder has produced g (gbuf) and dZ3 the rest of the backward pass is a DAG, not a
    // chain: {dW4, db4} | {dW3, db3} | dZ2 -> ({dW2, db2} | dZ1 -> ({dW1} | db1)). The branches are enqueued on the
    // organization's auxiliary streams between event forks/joins; under stream capture they become parallel branches
    // of the epoch graph, so a step's critical path is the dZ2 -> dZ1 -> dW1 chain instead of the sum of all kernels.
    // (`only` >= 0, the per-class profiler, keeps everything on the main stream.)
    const bool par = only < 0 && o->fanout != 0;
    cudaStream_t sA = par ? o->aux[0] : st, sB = par ? o->aux[1] : st, sC = par ? o->aux[2] : st;
    int ev_i = 0;
    auto fork = [&](cudaStream_t from, cudaStream_t to) -> int {  // `to` continues after everything enqueued on `from`
        if (from == to) return 0;
        cudaEvent_t ev = o->fev[ev_i++ % kForkEvents];
        DMT_CUDA(cudaEventRecord(ev, from));
        DMT_CUDA(cudaStreamWaitEvent(to, ev, 0));
        return 0;
    };
    if (only < 0 || only == K_SEG_W4 || only == K_DENSE_BWD) {
        if ((rc = fork(st, sA))) return rc;
        if ((rc = fork(st, sB))) return rc;
    }
    // dW4, db4: segmented over (batch, target column) — or the tensor-core G^T A3 — on branch A
    if (WANT(K_SEG_W4) && tc) {
        if ((rc = launch_decoder_tc_bwd_w(o->rows_buf, o->t_indptr, o->t_indices, o->gbuf, o->pt.ent_off, o->a3, H1,
                                          o->n_dec, G + o->oW4, G + o->ob4, tab, o->dec_passes, B, br, sA)))
            return rc;
    }
    if (WANT(K_SEG_W4) && !tc) {
        ChunkedSegs cs{o->pt.perm, o->pt.ent_row, o->pt.seg_key, o->pt.seg_off, o->pt.batch_seg_off,
                       o->pt.seg_chunk_off, o->pt.chunk_seg, o->pt.batch_chunk_off, o->pt.part, o->pt.part_bias, b,
                       o->n_dec};
        if ((rc = launch_segment_chunks(cs, o->n_dec * 2, o->n_dec, o->gbuf, o->a3, H1, G + o->oW4, G + o->ob4,
                                        o->active, sA)))
            return rc;
    }
    // dense backward: dW3/db3 on branch B, dZ2 on the main stream, then dW2/db2 on branch C next to dZ1
    if (WANT(K_DENSE_BWD)) {
        if ((rc = launch_dense_bwd_w(o->dz3, o->c, G + o->oW3, G + o->ob3, B, H1, H2, br, sB))) return rc;
        if ((rc = launch_dense_bwd_x(o->dz3, W3, o->a2, drop, o->dz2, B, H1, H2, 1, br, st))) return rc;
        if ((rc = fork(st, sC))) return rc;
        if ((rc = launch_dense_bwd_w(o->dz2, o->a1, G + o->oW2, G + o->ob2, B, H2, H1, br, sC))) return rc;
        if ((rc = launch_dense_bwd_x(o->dz2, W2, o->a1, nodrop, o->dz1, B, H2, H1, 1, br, st))) return rc;
    }
    // dW1t: segmented over (batch, data column) on the main stream; db1 = column sums of dZ1 on branch B
    if (WANT(K_SEG_W1)) {
        if ((rc = fork(st, sB))) return rc;
        if ((rc = launch_colsum(o->dz1, H1, G + o->ob1, br, sB))) return rc;
        ChunkedSegs cs{o->pd.perm, o->pd.ent_row, o->pd.seg_key, o->pd.seg_off, o->pd.batch_seg_off,
                       o->pd.seg_chunk_off, o->pd.chunk_seg, o->pd.batch_chunk_off, o->pd.part, o->pd.part_bias, b,
                       o->n_enc};
        if ((rc = launch_segment_chunks(cs, o->n_enc * 2, o->n_enc, o->dval_ord, o->dz1, H1, G + o->oW1, nullptr,
                                        o->active, st)))
            return rc;
    }
    // join: the norm reads every gradient
    if ((rc = fork(sA, st)) || (rc = fork(sB, st)) || (rc = fork(sC, st))) return rc;
    // clip + Adam
    if (WANT(K_NORM))
        if ((rc = launch_sqnorm_stage1(G, o->n_params, o->partial, br, st))) return rc;
    if (WANT(K_ADAM)) {
        if ((rc = launch_adam_prepare(o->partial, kNormBlocks, nullptr, nullptr, o->sc, hp, 0, o->step_dev,
                                      o->loss_rows, o->pt.batch_cnt + b, o->loss_buf + b, br, st)))
            return rc;
        if ((rc = launch_adam(o->P, G, o->M, o->V, o->n_params, o->sc, hp, true, st))) return rc;
    }
#undef WANT
    return 0;
}

// The table of a CSR (by CSR row id), built on first use on the organization's stream; null when it would be too large
// (the kernels then binary-search) or when all slots are taken.
static TcTab tile_tab_for(dmt_org* o, const int32_t* indptr, const int32_t* indices, int n_rows) {
    for (int i = 0; i < o->n_tabs; ++i)
        if (o->tabs[i].indptr == indptr && o->tabs[i].indices == indices && o->tabs[i].n_rows == n_rows)
            return TcTab{o->tabs[i].tab, 2};
    const int64_t n = decoder_tc_tab_ints(n_rows, o->n_dec);
    if (o->n_tabs >= 4 || n > (64LL << 20)) return TcTab{nullptr, 0};
    int32_t* tab = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&tab), (size_t)n * 4) != cudaSuccess) {
        cudaGetLastError();
        return TcTab{nullptr, 0};
    }
    if (build_tile_tab(nullptr, n_rows, indptr, indices, o->n_dec, tab, o->st)) {
        cudaFree(tab);
        return TcTab{nullptr, 0};
    }
    o->tabs[o->n_tabs++] = dmt_org::TabEntry{indptr, indices, n_rows, tab};
    return TcTab{tab, 2};
}

// W2t / W3t follow P: after every write of P that does not go through the fused Adam kernel
static int refresh_shadows(dmt_org* o) {
    if (o->W2t == nullptr) return 0;
    return launch_shadow_refresh(o->P + o->oW2, o->P + o->oW3, o->W2t, o->W3t, o->st);
}

static void drop_graph(dmt_org* o) {
    if (o->exec) {
        cudaGraphExecDestroy(o->exec);
        o->exec = nullptr;
    }
    o->g_nb = -1;
}

static bool same_hp(const AdamHyper& a, const AdamHyper& b) {
    return a.lr == b.lr && a.beta1 == b.beta1 && a.beta2 == b.beta2 && a.eps == b.eps &&
           a.weight_decay == b.weight_decay && a.max_norm == b.max_norm;
}

static int free_all(dmt_org* o) {
    if (o->exec) cudaGraphExecDestroy(o->exec);
    cudaFree(o->P); cudaFree(o->G); cudaFree(o->M); cudaFree(o->V);
    cudaFree(o->a1); cudaFree(o->a2); cudaFree(o->c); cudaFree(o->a3); cudaFree(o->dz3); cudaFree(o->dz2);
    cudaFree(o->dz1); cudaFree(o->loss_rows); cudaFree(o->iota_rows);
    free_side(o->pt); free_side(o->pd);
    cudaFree(o->gbuf); cudaFree(o->dval_ord); cudaFree(o->row_batch); cudaFree(o->active); cudaFree(o->sort_temp);
    cudaFree(o->t_nch_row); cudaFree(o->t_chunk_off); cudaFree(o->t_chunk_row); cudaFree(o->t_batch_chunk);
    cudaFree(o->dz_part); cudaFree(o->loss_part); cudaFree(o->tc_scratch);
    for (int i = 0; i < o->n_tabs; ++i) cudaFree(o->tabs[i].tab);
    cudaFree(o->rows_buf); cudaFree(o->row_off_buf); cudaFree(o->keep_buf); cudaFree(o->seed_dev);
    cudaFree(o->loss_buf); cudaFree(o->partial); cudaFree(o->sc); cudaFree(o->step_dev);
    cudaFree(o->W2t); cudaFree(o->W3t); cudaFree(o->dec_meta); cudaFree(o->row_cnt); cudaFree(o->g_sorted);
    cudaFree(o->t_seg_meta); cudaFree(o->d_seg_meta); cudaFree(o->t_row_sorted); cudaFree(o->d_row_sorted);
    cudaFree(o->t_inv_perm); cudaFree(o->d_val_sorted); cudaFree(o->t_seg_cnt); cudaFree(o->d_seg_cnt);
    cudaFree(o->part_db); cudaFree(o->part_dw); cudaFree(o->dw_cnt); cudaFree(o->norm_ticket);
    if (o->ev) cudaEventDestroy(o->ev);
    for (int i = 0; i < 3; ++i) if (o->aux[i]) cudaStreamDestroy(o->aux[i]);
    for (int i = 0; i < kForkEvents; ++i) if (o->fev[i]) cudaEventDestroy(o->fev[i]);
    if (o->own_stream && o->st) cudaStreamDestroy(o->st);
    return 0;
}

__global__ void iota32_kernel(int32_t* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace dmt

extern "C" {

int dmt_org_create(dmt_org_t** out, int n_rows, int n_enc, int n_dec, int H1, int H2, const int32_t* d_indptr,
                   const int32_t* d_indices, const float* d_val, int64_t d_nnz, const int32_t* t_indptr,
                   const int32_t* t_indices, int64_t t_nnz, int batch_rows, int loss_kind, int plan_epochs,
                   void* stream) {
    DMT_REQUIRE(out && n_rows > 0 && n_enc > 0 && n_dec > 0 && batch_rows > 0, "dmt_org_create: bad sizes");
    DMT_REQUIRE(H1 % 128 == 0 && H1 <= 512 && H2 > 0, "dmt_org_create: H1 must be 128/256/384/512");
    // the flat layout puts b2, W3, b3, W4 behind H2-sized blocks and every sub-view is read with 128-bit loads
    DMT_REQUIRE(H2 % 4 == 0, "dmt_org_create: H2 must be a multiple of 4");
    DMT_REQUIRE(d_nnz >= 0 && t_nnz >= 0 && t_nnz < (1LL << 31) - 2 && d_nnz < (1LL << 31) - 2,
                "dmt_org_create: nnz must fit int32");
    int rc = dmt_check_device();
    if (rc) return rc;
    dmt_org* o = new (std::nothrow) dmt_org();
    if (!o) return DMT_E_NOMEM;
    *o = dmt_org{};
    o->n_rows = n_rows; o->n_enc = n_enc; o->n_dec = n_dec; o->H1 = H1; o->H2 = H2;
    o->batch_rows = batch_rows; o->loss_kind = loss_kind;
    o->d_indptr = d_indptr; o->d_indices = d_indices; o->d_val = d_val; o->d_nnz = d_nnz;
    o->t_indptr = t_indptr; o->t_indices = t_indices; o->t_val = nullptr; o->t_nnz = t_nnz;
    if (stream) {
        o->st = as_stream(stream);
        o->own_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&o->st, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete o; set_error(cudaGetErrorString(e)); return (int)e; }
        o->own_stream = true;
    }
    int64_t off = 0;
    o->oW1 = off; off += (int64_t)n_enc * H1;
    o->ob1 = off; off += H1;
    o->oW2 = off; off += (int64_t)H2 * H1;
    o->ob2 = off; off += H2;
    o->oW3 = off; off += (int64_t)H1 * H2;
    o->ob3 = off; off += H1;
    o->oW4 = off; off += (int64_t)n_dec * H1;
    o->ob4 = off; off += n_dec;
    o->n_params = off;
    o->act_rows = n_rows < 65536 ? n_rows : 65536;
    if (o->act_rows < batch_rows) o->act_rows = batch_rows;
    // one plan / one graph may cover several local epochs at once (plan_epochs): the plan buffers are sized for it
    if (plan_epochs < 1) plan_epochs = 1;
    DMT_REQUIRE((int64_t)t_nnz * plan_epochs < (1LL << 31) - 2 && (int64_t)d_nnz * plan_epochs < (1LL << 31) - 2,
                "dmt_org_create: plan_epochs x nnz must fit int32");
    const int64_t t_cap = t_nnz * plan_epochs, d_cap = d_nnz * plan_epochs;
    o->rows_cap = n_rows * plan_epochs;
    o->nb_cap = ((n_rows + batch_rows - 1) / batch_rows + 1) * plan_epochs;
#define A(expr) if ((rc = (expr))) { free_all(o); delete o; return rc; }
    A(dalloc(&o->P, o->n_params)); A(dalloc(&o->G, o->n_params)); A(dalloc(&o->M, o->n_params));
    A(dalloc(&o->V, o->n_params));
    int64_t ar = o->act_rows;
    A(dalloc(&o->a1, ar * H1)); A(dalloc(&o->a2, ar * H2)); A(dalloc(&o->c, ar * H2)); A(dalloc(&o->a3, ar * H1));
    A(dalloc(&o->dz3, (int64_t)batch_rows * H1)); A(dalloc(&o->dz2, (int64_t)batch_rows * H2));
    A(dalloc(&o->dz1, (int64_t)batch_rows * H1)); A(dalloc(&o->loss_rows, batch_rows));
    A(dalloc(&o->iota_rows, n_rows));
    A(alloc_side(o->pt, t_cap, o->rows_cap, o->nb_cap, n_dec, H1, batch_rows));
    A(alloc_side(o->pd, d_cap, o->rows_cap, o->nb_cap, n_enc, H1, batch_rows));
    o->dec_chunk_cap = t_cap / kDecChunk + o->rows_cap + 16;
    {
        int64_t per_batch = (int64_t)batch_rows * n_dec < t_cap ? (int64_t)batch_rows * n_dec : t_cap;
        o->dec_part_rows = per_batch / kDecChunk + batch_rows + 16;
    }
    A(dalloc(&o->t_nch_row, o->rows_cap + 2)); A(dalloc(&o->t_chunk_off, o->rows_cap + 2));
    A(dalloc(&o->t_chunk_row, o->dec_chunk_cap)); A(dalloc(&o->t_batch_chunk, o->nb_cap + 2));
    A(dalloc(&o->dz_part, o->dec_part_rows * H1)); A(dalloc(&o->loss_part, o->dec_part_rows));
    o->dec_mode = 0; o->dec_passes = 3;
    {
        const char* env = getenv("DMT_DEC_BLOCKS");
        o->dec_blocks = env ? atoi(env) : 0;
        env = getenv("DMT_GATHER");
        o->gather = (env && strcmp(env, "bulk") == 0) ? 1 : 0;  // measured: per-row bulk copies are slower
        env = getenv("DMT_BULK_BLOCKS");
        o->bulk_blocks = env ? atoi(env) : 0;
        env = getenv("DMT_PDL");
        o->pdl = (env && atoi(env) != 0) ? 1 : 0;
        env = getenv("DMT_DEC_FORM");
        o->dec_form = (env && strcmp(env, "8w") == 0) ? 0 : 2;
        // Row kernels: weights through shared memory by bulk copies (fused_rows.cu) for small batches, where a row
        // tile of 2-4 rows and several warps per row pay (Douban shape, 100-row batches: forward rows 130 -> 37 us,
        // round 240 -> 190 ms); at 500-row batches both forms take 19 / 16 us per launch and the register-streamed
        // form packs better next to other organizations' kernels (18 organizations: 198 against 224 ms per round).
        env = getenv("DMT_ROWS");
        o->rows_R = fused_rows_per_cta(batch_rows);
        o->rows_mode = o->rows_R < 8 ? 1 : 0;
        if (env) o->rows_mode = strcmp(env, "tma") == 0 ? 1 : 0;
        env = getenv("DMT_STREAM_ROWS");
        o->stream_R = env ? atoi(env) : kFusedRows;
        if (o->stream_R != 4 && o->stream_R != 16) o->stream_R = kFusedRows;
    }
    A(dalloc(&o->tc_scratch, decoder_tc_scratch_floats(batch_rows, n_dec, H1)));
    A(dalloc(&o->gbuf, t_cap)); A(dalloc(&o->dval_ord, d_cap)); A(dalloc(&o->row_batch, o->rows_cap + 1));
    A(dalloc(&o->active, o->nb_cap + 1));
    o->sort_temp_bytes = sort_segments_temp_bytes(t_cap > d_cap ? t_cap : d_cap);
    if (o->sort_temp_bytes < (1 << 20)) o->sort_temp_bytes = 1 << 20;
    A(dalloc(reinterpret_cast<char**>(&o->sort_temp), o->sort_temp_bytes));
    A(dalloc(&o->rows_buf, o->rows_cap)); A(dalloc(&o->row_off_buf, o->nb_cap + 1));
    o->keep_buf = nullptr;  // explicit dropout masks are a parity-test input: allocated on first use (train_epoch)
    A(dalloc(&o->seed_dev, 1)); A(dalloc(&o->loss_buf, o->nb_cap));
    A(dalloc(&o->partial, kNormBlocks)); A(dalloc(&o->sc, 1)); A(dalloc(&o->step_dev, 1));
    // fused step: usable for the reference AE shape; batch_rows <= 512 keeps the dW tiles at <= 8 row slices
    o->step_mode = (H1 == 256 && H2 == 128 && batch_rows <= 8 * kDwSlice &&
                    (int64_t)n_dec <= (int64_t)kFusedMaxRowChunks * kDecChunk) ? 1 : 0;
    {
        const char* env = getenv("DMT_STEP");
        if (env && strcmp(env, "classic") == 0) o->step_mode = 0;
    }
    if (H1 == 256 && H2 == 128) {
        A(dalloc(&o->W2t, (int64_t)H1 * H2)); A(dalloc(&o->W3t, (int64_t)H1 * H2));
        A(dalloc(&o->dec_meta, o->dec_chunk_cap)); A(dalloc(&o->row_cnt, batch_rows)); A(dalloc(&o->g_sorted, t_cap));
        A(dalloc(&o->t_seg_meta, o->pt.chunk_cap)); A(dalloc(&o->d_seg_meta, o->pd.chunk_cap));
        A(dalloc(&o->t_row_sorted, t_cap)); A(dalloc(&o->d_row_sorted, d_cap)); A(dalloc(&o->t_inv_perm, t_cap));
        A(dalloc(&o->d_val_sorted, d_cap));
        A(dalloc(&o->t_seg_cnt, o->pt.part_rows)); A(dalloc(&o->d_seg_cnt, o->pd.part_rows));
        A(dalloc(&o->part_db, (int64_t)((batch_rows + 1) / 2 + 1) * kDbPartStride));  // row tiles of >= 2 rows
        A(prepare_fused_rows());
        A(dalloc(&o->part_dw, (int64_t)16 * 8 * 4096)); A(dalloc(&o->dw_cnt, 16)); A(dalloc(&o->norm_ticket, 1));
        cudaMemsetAsync(o->row_cnt, 0, (size_t)batch_rows * 4, o->st);
        cudaMemsetAsync(o->t_seg_cnt, 0, (size_t)o->pt.part_rows * 4, o->st);
        cudaMemsetAsync(o->d_seg_cnt, 0, (size_t)o->pd.part_rows * 4, o->st);
        cudaMemsetAsync(o->dw_cnt, 0, 16 * 4, o->st);
        cudaMemsetAsync(o->norm_ticket, 0, 4, o->st);
    }
#undef A
    {
        cudaError_t e = cudaEventCreateWithFlags(&o->ev, cudaEventDisableTiming);
        for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&o->aux[i], cudaStreamNonBlocking);
        for (int i = 0; i < kForkEvents && e == cudaSuccess; ++i)
            e = cudaEventCreateWithFlags(&o->fev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { free_all(o); delete o; set_error(cudaGetErrorString(e)); return (int)e; }
    }
    iota32_kernel<<<(n_rows + 255) / 256, 256, 0, o->st>>>(o->iota_rows, n_rows);
    cudaMemsetAsync(o->step_dev, 0, sizeof(int), o->st);
    cudaMemsetAsync(o->M, 0, (size_t)o->n_params * 4, o->st);
    cudaMemsetAsync(o->V, 0, (size_t)o->n_params * 4, o->st);
    cudaMemsetAsync(o->G, 0, (size_t)o->n_params * 4, o->st);
    o->exec = nullptr;
    o->g_nb = -1;
    *out = o;
    return 0;
}

int dmt_org_destroy(dmt_org_t* o) {
    if (!o) return 0;
    cudaStreamSynchronize(o->st);
    free_all(o);
    delete o;
    return 0;
}

int64_t dmt_org_num_params(const dmt_org_t* o) { return o ? o->n_params : 0; }

int dmt_org_set_decoder_blocks(dmt_org_t* o, int blocks) {
    DMT_REQUIRE(o && blocks >= 0 && blocks <= 65535, "dmt_org_set_decoder_blocks: bad argument");
    if (o->dec_blocks != blocks && o->exec) {  // the grid is baked into the captured graph
        cudaGraphExecDestroy(o->exec);
        o->exec = nullptr;
        o->g_nb = -1;
    }
    o->dec_blocks = blocks;
    return 0;
}

int dmt_org_set_gather_mode(dmt_org_t* o, int mode) {
    DMT_REQUIRE(o && (mode == 0 || mode == 1), "dmt_org_set_gather_mode: bad argument");
    if (o->gather != mode) drop_graph(o);  // the kernels are baked into the captured graph
    o->gather = mode;
    return 0;
}

int dmt_org_gather_mode(const dmt_org_t* o) { return o ? o->gather : 0; }

int dmt_org_set_row_tile(dmt_org_t* o, int rows) {
    DMT_REQUIRE(o && (rows == 4 || rows == 8 || rows == 16), "dmt_org_set_row_tile: rows must be 4, 8 or 16");
    if (getenv("DMT_STREAM_ROWS") != nullptr) return 0;  // an explicit environment choice wins (A/B runs)
    if (o->stream_R != rows) drop_graph(o);  // grid and kernel instance are baked into the captured graph
    o->stream_R = rows;
    return 0;
}

int dmt_org_set_pdl(dmt_org_t* o, int on) {
    DMT_REQUIRE(o, "dmt_org_set_pdl: null");
    on = on ? 1 : 0;
    if (o->pdl != on) drop_graph(o);  // the launch attribute is baked into the captured graph
    o->pdl = on;
    return 0;
}

int dmt_org_set_fanout(dmt_org_t* o, int on) {
    DMT_REQUIRE(o, "dmt_org_set_fanout: null");
    on = on ? 1 : 0;
    if (o->fanout != on && o->exec) {  // the shape of the graph changes
        cudaGraphExecDestroy(o->exec);
        o->exec = nullptr;
        o->g_nb = -1;
    }
    o->fanout = on;
    return 0;
}

int dmt_org_set_decoder_mode(dmt_org_t* o, int mode, int passes) {
    DMT_REQUIRE(o && (mode == 0 || mode == 1) && (passes == 1 || passes == 3), "dmt_org_set_decoder_mode: bad argument");
    if ((o->dec_mode != mode || o->dec_passes != passes) && o->exec) {  // the choice is baked into the captured graph
        cudaGraphExecDestroy(o->exec);
        o->exec = nullptr;
        o->g_nb = -1;
    }
    const bool to_fused = o->step_mode == 1 && o->dec_mode == 1 && mode == 0;
    o->dec_mode = mode;
    o->dec_passes = passes;
    if (mode == 1 && o->train_tab.tab == nullptr) o->train_tab = tile_tab_for(o, o->t_indptr, o->t_indices, o->n_rows);
    if (to_fused) return refresh_shadows(o);  // the classic step's Adam does not maintain them
    return 0;
}

int dmt_org_set_step_mode(dmt_org_t* o, int mode) {
    DMT_REQUIRE(o && (mode == 0 || mode == 1), "dmt_org_set_step_mode: bad argument");
    DMT_REQUIRE(mode == 0 || (o->W2t != nullptr && o->batch_rows <= 8 * kDwSlice &&
                              (int64_t)o->n_dec <= (int64_t)kFusedMaxRowChunks * kDecChunk),
                "dmt_org_set_step_mode: the fused step needs H1 = 256, H2 = 128, batch_rows <= 512");
    if (o->step_mode == mode) return 0;
    drop_graph(o);  // the shape of the graph changes; plans of the fused step carry extra tables
    o->step_mode = mode;
    if (mode == 1) return refresh_shadows(o);
    return 0;
}

int dmt_org_step_mode(const dmt_org_t* o) { return o ? (o->step_mode == 1 && o->dec_mode == 0 ? 1 : 0) : 0; }

int dmt_org_set_params(dmt_org_t* o, const float* flat) {
    DMT_REQUIRE(o && flat, "dmt_org_set_params: null");
    DMT_CUDA(cudaMemcpyAsync(o->P, flat, (size_t)o->n_params * 4, cudaMemcpyDeviceToDevice, o->st));
    DMT_CUDA(cudaMemsetAsync(o->M, 0, (size_t)o->n_params * 4, o->st));
    DMT_CUDA(cudaMemsetAsync(o->V, 0, (size_t)o->n_params * 4, o->st));
    DMT_CUDA(cudaMemsetAsync(o->G, 0, (size_t)o->n_params * 4, o->st));
    DMT_CUDA(cudaMemsetAsync(o->step_dev, 0, sizeof(int), o->st));
    return refresh_shadows(o);
}

int dmt_org_get_params(const dmt_org_t* o, float* flat) {
    DMT_REQUIRE(o && flat, "dmt_org_get_params: null");
    DMT_CUDA(cudaMemcpyAsync(flat, o->P, (size_t)o->n_params * 4, cudaMemcpyDeviceToDevice, o->st));
    return 0;
}

int dmt_org_set_target(dmt_org_t* o, const float* t_val) {
    DMT_REQUIRE(o && t_val, "dmt_org_set_target: null");
    if (o->t_val != t_val && o->exec) {  // the pointer is baked into the captured graph
        cudaGraphExecDestroy(o->exec);
        o->exec = nullptr;
        o->g_nb = -1;
    }
    o->t_val = t_val;
    return 0;
}

int dmt_org_train_epoch(dmt_org_t* o, const int32_t* rows, const int32_t* row_off, int n_rows_total, int n_batches,
                        int64_t n_t_entries, int64_t n_d_entries, const uint8_t* keep, uint64_t seed, double lr,
                        double beta1, double beta2, double eps, double weight_decay, float max_norm,
                        float* epoch_loss) {
    DMT_REQUIRE(o && rows && row_off, "dmt_org_train_epoch: null");
    DMT_REQUIRE(o->t_val != nullptr, "dmt_org_train_epoch: call dmt_org_set_target first");
    DMT_REQUIRE(n_batches >= 1 && n_batches <= o->nb_cap && n_rows_total >= 0 && n_rows_total <= o->rows_cap,
                "dmt_org_train_epoch: too many rows or batches");
    DMT_REQUIRE(n_t_entries <= o->pt.cap && n_d_entries <= o->pd.cap, "dmt_org_train_epoch: entry count over capacity");
    DMT_REQUIRE((int64_t)n_batches * (o->n_dec > o->n_enc ? o->n_dec : o->n_enc) < (1LL << 32),
                "dmt_org_train_epoch: batches x columns must fit 32-bit sort keys");
    cudaStream_t st = o->st;
    DMT_CUDA(cudaMemcpyAsync(o->rows_buf, rows, (size_t)n_rows_total * 4, cudaMemcpyDeviceToDevice, st));
    DMT_CUDA(cudaMemcpyAsync(o->row_off_buf, row_off, (size_t)(n_batches + 1) * 4, cudaMemcpyDeviceToDevice, st));
    if (keep) {
        if (!o->keep_buf) {
            int rc_k = dalloc(&o->keep_buf, (int64_t)o->rows_cap * o->H2);
            if (rc_k) return rc_k;
            if (o->exec) { cudaGraphExecDestroy(o->exec); o->exec = nullptr; o->g_nb = -1; }  // pointer is baked in
        }
        DMT_CUDA(cudaMemcpyAsync(o->keep_buf, keep, (size_t)n_rows_total * o->H2, cudaMemcpyDeviceToDevice, st));
    }
    DMT_CUDA(cudaMemcpyAsync(o->seed_dev, &seed, sizeof(seed), cudaMemcpyHostToDevice, st));
    int rc = 0;
    AdamHyper hp{lr, beta1, beta2, eps, weight_decay, max_norm};
    int use_keep = keep != nullptr;
    // The whole epoch — plan (scans, the two radix sorts, segment/chunk tables) AND the per-batch steps — is one graph:
    // one launch per epoch from the host. Entry totals are baked into the captured sort calls, so the graph is keyed
    // on them too (they are constant for epochs that cover every row).
    if (!o->exec || o->g_nb != n_batches || o->g_keep != use_keep || !same_hp(hp, o->g_hp) ||
        o->g_rows != n_rows_total || o->g_nt != n_t_entries || o->g_nd != n_d_entries) {
        if (o->exec) { cudaGraphExecDestroy(o->exec); o->exec = nullptr; }
        cudaGraph_t graph = nullptr;
        DMT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        long long before = launch_count();
        rc = build_plan(o, n_rows_total, n_batches, n_t_entries, n_d_entries, st);
        for (int b = 0; b < n_batches && rc == 0; ++b) rc = enqueue_step(o, b, use_keep != 0, hp);
        o->g_kernels = launch_count() - before;
        count_launch(-o->g_kernels);  // captured, not executed: counted at every cudaGraphLaunch instead
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        e = cudaGraphInstantiate(&o->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { o->exec = nullptr; set_error(cudaGetErrorString(e)); return (int)e; }
        o->g_nb = n_batches; o->g_keep = use_keep; o->g_hp = hp;
        o->g_rows = n_rows_total; o->g_nt = n_t_entries; o->g_nd = n_d_entries;
    }
    DMT_CUDA(cudaGraphLaunch(o->exec, st));
    count_launch(o->g_kernels);
    if (epoch_loss)
        DMT_CUDA(cudaMemcpyAsync(epoch_loss, o->loss_buf, (size_t)n_batches * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int dmt_org_predict(dmt_org_t* o, const int32_t* d_indptr, const int32_t* d_indices, const float* d_val,
                    const int32_t* t_indptr, const int32_t* t_indices, int n_rows, float* pred) {
    DMT_REQUIRE(o && pred && n_rows >= 0 && n_rows <= o->n_rows, "dmt_org_predict: bad argument");
    cudaStream_t st = o->st;
    const int H1 = o->H1, H2 = o->H2;
    float *W1t = o->P + o->oW1, *b1 = o->P + o->ob1, *W2 = o->P + o->oW2, *b2 = o->P + o->ob2;
    float *W3 = o->P + o->oW3, *b3 = o->P + o->ob3, *W4 = o->P + o->oW4, *b4 = o->P + o->ob4;
    Dropout nodrop;
    int rc;
    for (int lo = 0; lo < n_rows; lo += o->act_rows) {
        int hi = lo + o->act_rows < n_rows ? lo + o->act_rows : n_rows;
        BatchRef br = batch_by_value(lo, hi);
        int m = hi - lo;
        if (o->step_mode == 1 && o->dec_mode == 0) {  // encoder + both dense layers in one row-local launch
            FusedFwd f{br, o->iota_rows, d_indptr, d_indices, d_val, W1t, b1, o->W2t, b2, o->W3t, b3,
                       nullptr, nullptr, nullptr, o->a3, nodrop};
            if ((rc = o->rows_mode ? launch_fused_fwd_tma(f, m, fused_rows_per_cta(m), st) : launch_fused_fwd(f, m, st)))
                return rc;
        } else {
            if ((rc = launch_ae_encoder_fwd(o->iota_rows, d_indptr, d_indices, d_val, W1t, b1, H1, o->a1, m, br, st)))
                return rc;
            if ((rc = launch_dense_fwd(o->a1, W2, b2, o->c, nullptr, nodrop, m, H2, H1, 1, br, st))) return rc;
            if ((rc = launch_dense_fwd(o->c, W3, b3, o->a3, nullptr, nodrop, m, H1, H2, 1, br, st))) return rc;
        }
        if (o->dec_mode == 1) {
            // rows [lo, hi) of the split: a3 holds them from row 0, the CSR rows are iota_rows[lo + j]
            if ((rc = launch_decoder_tc_fwd(o->iota_rows, t_indptr, t_indices, nullptr, o->a3, W4, b4, H1, o->n_dec,
                                            o->loss_kind, nullptr, nullptr, pred, nullptr, nullptr,
                                            tile_tab_for(o, t_indptr, t_indices, n_rows), o->dec_passes, m, br, st)))
                return rc;
        } else if ((rc = launch_ae_decoder_fwd(o->iota_rows, t_indptr, t_indices, nullptr, o->a3, W4, b4, H1,
                                               o->loss_kind, nullptr, nullptr, pred, nullptr, nullptr, nullptr, 0, m,
                                               br, st)))
            return rc;
    }
    return 0;
}

/* Average duration (ms) of every kernel class of one training step on batch b of the CURRENT plan (call after
 * dmt_org_train_epoch): each class is launched `reps` times back to back between two CUDA events on the handle's
 * stream. Parameters and optimizer state are restored afterwards. ms_out has dmt_org_profile_classes() entries:
 * encoder, dense fwd (2 GEMMs), grad zero, decoder+loss+dZ3, dW4 segments, dense bwd (4 GEMMs + 2 col-sums),
 * dW1 segments + col-sum, grad norm, Adam. */
int dmt_org_profile_classes(void) { return K_NCLASS; }

int dmt_org_profile_step(dmt_org_t* o, int b, int reps, float* ms_out) {
    DMT_REQUIRE(o && ms_out && reps > 0 && b >= 0 && b < o->nb_cap, "dmt_org_profile_step: bad argument");
    cudaStream_t st = o->st;
    float *P0 = nullptr, *M0 = nullptr, *V0 = nullptr;
    int step0 = 0;
    size_t bytes = (size_t)o->n_params * 4;
    DMT_CUDA(cudaMalloc(&P0, bytes)); DMT_CUDA(cudaMalloc(&M0, bytes)); DMT_CUDA(cudaMalloc(&V0, bytes));
    DMT_CUDA(cudaMemcpyAsync(P0, o->P, bytes, cudaMemcpyDeviceToDevice, st));
    DMT_CUDA(cudaMemcpyAsync(M0, o->M, bytes, cudaMemcpyDeviceToDevice, st));
    DMT_CUDA(cudaMemcpyAsync(V0, o->V, bytes, cudaMemcpyDeviceToDevice, st));
    DMT_CUDA(cudaMemcpyAsync(&step0, o->step_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    cudaEvent_t e0, e1;
    DMT_CUDA(cudaEventCreate(&e0)); DMT_CUDA(cudaEventCreate(&e1));
    AdamHyper hp = o->g_nb >= 0 ? o->g_hp : AdamHyper{1e-3, 0.9, 0.999, 1e-8, 5e-4, 1.f};
    int rc = 0;
    // one full step first so every buffer a class reads holds this batch's data
    rc = enqueue_step(o, b, o->g_keep != 0, hp);
    for (int cls = 0; cls < K_NCLASS && rc == 0; ++cls) {
        rc = enqueue_step(o, b, o->g_keep != 0, hp, cls);  // warm
        DMT_CUDA(cudaEventRecord(e0, st));
        for (int r = 0; r < reps && rc == 0; ++r) rc = enqueue_step(o, b, o->g_keep != 0, hp, cls);
        DMT_CUDA(cudaEventRecord(e1, st));
        DMT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        DMT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        ms_out[cls] = ms / reps;
    }
    cudaMemcpyAsync(o->P, P0, bytes, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(o->M, M0, bytes, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(o->V, V0, bytes, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(o->step_dev, &step0, sizeof(int), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(P0); cudaFree(M0); cudaFree(V0);
    return rc;
}

int dmt_org_sync(dmt_org_t* o) {
    DMT_REQUIRE(o, "dmt_org_sync: null");
    DMT_CUDA(cudaStreamSynchronize(o->st));
    return 0;
}

void* dmt_org_stream(dmt_org_t* o) { return o ? (void*)o->st : nullptr; }

int dmt_org_wait_stream(dmt_org_t* o, void* stream) {
    DMT_REQUIRE(o, "dmt_org_wait_stream: null");
    if (as_stream(stream) == o->st) return 0;
    DMT_CUDA(cudaEventRecord(o->ev, as_stream(stream)));
    DMT_CUDA(cudaStreamWaitEvent(o->st, o->ev, 0));
    return 0;
}

int dmt_org_signal_stream(dmt_org_t* o, void* stream) {
    DMT_REQUIRE(o, "dmt_org_signal_stream: null");
    if (as_stream(stream) == o->st) return 0;
    DMT_CUDA(cudaEventRecord(o->ev, o->st));
    DMT_CUDA(cudaStreamWaitEvent(as_stream(stream), o->ev, 0));
    return 0;
}


/* ---------------------------------------------------------------- organization groups */

struct dmt_group {
    std::vector<dmt_org*> orgs;
    std::vector<OrgDev> views;  // what d_orgs currently holds
    OrgDev* d_orgs;
    cudaStream_t st;
    bool own_stream;
    cudaEvent_t ev;
    cudaGraphExec_t exec;
    long long g_kernels;
    int g_nb;
    AdamHyper g_hp;
    int64_t n_params_max;
    int n_enc_max;
};

static void group_free(dmt_group* g) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->d_orgs) cudaFree(g->d_orgs);
    if (g->ev) cudaEventDestroy(g->ev);
    if (g->own_stream && g->st) cudaStreamDestroy(g->st);
    delete g;
}

static OrgDev org_view(dmt_org* o) {
    OrgDev v{};
    v.rows = o->rows_buf; v.row_off = o->row_off_buf; v.active = o->active;
    v.d_indptr = o->d_indptr; v.d_indices = o->d_indices; v.t_indptr = o->t_indptr; v.t_indices = o->t_indices;
    v.d_val = o->d_val; v.t_val = o->t_val;
    v.P = o->P; v.G = o->G; v.M = o->M; v.V = o->V; v.n_params = o->n_params;
    v.oW1 = o->oW1; v.ob1 = o->ob1; v.oW2 = o->oW2; v.ob2 = o->ob2; v.oW3 = o->oW3; v.ob3 = o->ob3;
    v.oW4 = o->oW4; v.ob4 = o->ob4; v.n_enc = o->n_enc; v.n_dec = o->n_dec;
    v.a1 = o->a1; v.a2 = o->a2; v.c = o->c; v.a3 = o->a3; v.dz3 = o->dz3; v.dz2 = o->dz2; v.dz1 = o->dz1;
    v.loss_rows = o->loss_rows;
    v.t_ent_off = o->pt.ent_off; v.t_batch_cnt = o->pt.batch_cnt;
    v.dc = DecChunks{o->t_chunk_off, o->t_chunk_row, o->t_batch_chunk, o->dz_part, o->loss_part};
    v.gbuf = o->gbuf;
    v.seg_t = ChunkedSegs{o->pt.perm, o->pt.ent_row, o->pt.seg_key, o->pt.seg_off, o->pt.batch_seg_off,
                          o->pt.seg_chunk_off, o->pt.chunk_seg, o->pt.batch_chunk_off, o->pt.part, o->pt.part_bias, 0,
                          o->n_dec};
    v.seg_d = ChunkedSegs{o->pd.perm, o->pd.ent_row, o->pd.seg_key, o->pd.seg_off, o->pd.batch_seg_off,
                          o->pd.seg_chunk_off, o->pd.chunk_seg, o->pd.batch_chunk_off, o->pd.part, o->pd.part_bias, 0,
                          o->n_enc};
    v.dval_ord = o->dval_ord;
    v.partial = o->partial; v.sc = o->sc; v.step_dev = o->step_dev; v.loss_buf = o->loss_buf; v.seed_dev = o->seed_dev;
    return v;
}

// The graph only bakes the address of the view array, so new target pointers etc. need no re-capture: the views are
// re-uploaded whenever they differ from what the device holds.
static int upload_views(dmt_group* g) {
    std::vector<OrgDev> host;
    for (dmt_org* o : g->orgs) host.push_back(org_view(o));
    if (g->views.size() == host.size() && memcmp(g->views.data(), host.data(), host.size() * sizeof(OrgDev)) == 0)
        return 0;
    DMT_CUDA(cudaStreamSynchronize(g->st));  // nothing in flight may still read the old views
    g->views = host;
    DMT_CUDA(cudaMemcpy(g->d_orgs, g->views.data(), host.size() * sizeof(OrgDev), cudaMemcpyHostToDevice));
    return 0;
}

static int enqueue_group_step(dmt_group* g, int b, AdamHyper hp) {
    dmt_org* o0 = g->orgs[0];
    const int G = (int)g->orgs.size(), B = o0->batch_rows, H1 = o0->H1, H2 = o0->H2;
    cudaStream_t st = g->st;
    int rc;
    if ((rc = launch_group_encoder(g->d_orgs, G, b, B, H1, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 0, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 1, st))) return rc;
    if ((rc = launch_group_decoder(g->d_orgs, G, b, B, H1, st))) return rc;
    if ((rc = launch_group_segments(g->d_orgs, G, b, 0, o0->n_dec, H1, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 2, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 3, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 4, st))) return rc;
    if ((rc = launch_group_dense(g->d_orgs, G, b, B, H1, H2, 5, st))) return rc;
    if ((rc = launch_group_segments(g->d_orgs, G, b, 1, g->n_enc_max, H1, st))) return rc;
    if ((rc = launch_group_colsum_dz1(g->d_orgs, G, b, H1, st))) return rc;
    if ((rc = launch_group_optim(g->d_orgs, G, b, g->n_params_max, hp, st))) return rc;
    return 0;
}

int dmt_group_create(dmt_group_t** out, dmt_org_t* const* orgs, int n, void* stream) {
    DMT_REQUIRE(out && orgs && n >= 1, "dmt_group_create: bad argument");
    dmt_group* g = new (std::nothrow) dmt_group();
    if (!g) return DMT_E_NOMEM;
    g->exec = nullptr; g->g_nb = -1; g->d_orgs = nullptr; g->n_params_max = 0; g->n_enc_max = 0;
    g->st = nullptr; g->own_stream = false; g->ev = nullptr;
    for (int i = 0; i < n; ++i) {
        dmt_org* o = orgs[i];
        if (!o || o->n_rows != orgs[0]->n_rows || o->H1 != orgs[0]->H1 || o->H2 != orgs[0]->H2 ||
            o->batch_rows != orgs[0]->batch_rows || o->n_dec != orgs[0]->n_dec) {
            delete g;
            set_error("dmt_group_create: organizations must share rows, hidden sizes, batch size and target width");
            return DMT_E_ARG;
        }
        g->orgs.push_back(o);
        if (o->n_params > g->n_params_max) g->n_params_max = o->n_params;
        if (o->n_enc > g->n_enc_max) g->n_enc_max = o->n_enc;
    }
    cudaError_t e = cudaSuccess;
    if (stream) g->st = as_stream(stream);
    else if ((e = cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking)) == cudaSuccess) g->own_stream = true;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->d_orgs), sizeof(OrgDev) * n);
    if (e != cudaSuccess) {
        group_free(g);
        set_error(cudaGetErrorString(e));
        return (int)e;
    }
    *out = g;
    return 0;
}

int dmt_group_destroy(dmt_group_t* g) {
    if (!g) return 0;
    cudaStreamSynchronize(g->st);
    group_free(g);
    return 0;
}

void* dmt_group_stream(dmt_group_t* g) { return g ? (void*)g->st : nullptr; }

int dmt_group_train(dmt_group_t* g, const int32_t* const* rows, const int32_t* const* row_off, int n_rows_total,
                    int n_batches, const int64_t* n_t_entries, const int64_t* n_d_entries, const uint64_t* seeds,
                    double lr, double beta1, double beta2, double eps, double weight_decay, float max_norm,
                    float* const* batch_loss) {
    DMT_REQUIRE(g && rows && row_off && n_t_entries && n_d_entries && seeds, "dmt_group_train: null");
    const int G = (int)g->orgs.size();
    AdamHyper hp{lr, beta1, beta2, eps, weight_decay, max_norm};
    int rc = 0;
    // 1. per-organization plans, each on the organization's own stream (they are independent), joined below
    for (int i = 0; i < G; ++i) {
        dmt_org* o = g->orgs[i];
        DMT_REQUIRE(o->t_val != nullptr, "dmt_group_train: call dmt_org_set_target on every organization first");
        DMT_REQUIRE(n_batches >= 1 && n_batches <= o->nb_cap && n_rows_total <= o->rows_cap &&
                        n_t_entries[i] <= o->pt.cap && n_d_entries[i] <= o->pd.cap,
                    "dmt_group_train: over the plan capacity (raise plan_epochs at dmt_org_create)");
        DMT_REQUIRE((int64_t)n_batches * (o->n_dec > o->n_enc ? o->n_dec : o->n_enc) < (1LL << 32),
                    "dmt_group_train: batches x columns must fit 32-bit sort keys");
        cudaStream_t st = o->st;
        DMT_CUDA(cudaEventRecord(g->ev, g->st));  // order after whatever the caller enqueued on the group stream
        DMT_CUDA(cudaStreamWaitEvent(st, g->ev, 0));
        DMT_CUDA(cudaMemcpyAsync(o->rows_buf, rows[i], (size_t)n_rows_total * 4, cudaMemcpyDeviceToDevice, st));
        DMT_CUDA(cudaMemcpyAsync(o->row_off_buf, row_off[i], (size_t)(n_batches + 1) * 4, cudaMemcpyDeviceToDevice, st));
        DMT_CUDA(cudaMemcpyAsync(o->seed_dev, &seeds[i], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if ((rc = build_plan(o, n_rows_total, n_batches, n_t_entries[i], n_d_entries[i], st))) return rc;
        DMT_CUDA(cudaEventRecord(o->ev, st));
        DMT_CUDA(cudaStreamWaitEvent(g->st, o->ev, 0));
    }
    // 2. one graph for all steps of all organizations
    if ((rc = upload_views(g))) return rc;
    bool stale = !g->exec || g->g_nb != n_batches || !same_hp(hp, g->g_hp);
    if (stale) {
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
        cudaGraph_t graph = nullptr;
        DMT_CUDA(cudaStreamBeginCapture(g->st, cudaStreamCaptureModeThreadLocal));
        long long before = launch_count();
        for (int b = 0; b < n_batches && rc == 0; ++b) rc = enqueue_group_step(g, b, hp);
        g->g_kernels = launch_count() - before;
        count_launch(-g->g_kernels);
        cudaError_t e = cudaStreamEndCapture(g->st, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
        e = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { g->exec = nullptr; set_error(cudaGetErrorString(e)); return (int)e; }
        g->g_nb = n_batches; g->g_hp = hp;
    }
    DMT_CUDA(cudaGraphLaunch(g->exec, g->st));
    count_launch(g->g_kernels);
    if (batch_loss)
        for (int i = 0; i < G; ++i)
            if (batch_loss[i])
                DMT_CUDA(cudaMemcpyAsync(batch_loss[i], g->orgs[i]->loss_buf, (size_t)n_batches * 4,
                                         cudaMemcpyDeviceToDevice, g->st));
    // later per-organization work (predict, get_params) must see the trained parameters
    DMT_CUDA(cudaEventRecord(g->ev, g->st));
    for (int i = 0; i < G; ++i) {
        DMT_CUDA(cudaStreamWaitEvent(g->orgs[i]->st, g->ev, 0));
        if ((rc = refresh_shadows(g->orgs[i]))) return rc;  // the group's Adam wrote P
    }
    return 0;
}

int dmt_group_sync(dmt_group_t* g) {
    DMT_REQUIRE(g, "dmt_group_sync: null");
    DMT_CUDA(cudaStreamSynchronize(g->st));
    return 0;
}

int dmt_group_wait_stream(dmt_group_t* g, void* stream) {
    DMT_REQUIRE(g, "dmt_group_wait_stream: null");
    if (as_stream(stream) == g->st) return 0;
    DMT_CUDA(cudaEventRecord(g->ev, as_stream(stream)));
    DMT_CUDA(cudaStreamWaitEvent(g->st, g->ev, 0));
    return 0;
}

}  // extern "C"
