// Internal launch functions shared by the stateless C-ABI wrappers and the organization engine.
#pragma once
#include "common.cuh"

namespace dmt {

constexpr int kNormBlocks = kNumSMs * 2;

struct AdamHyper {
    double lr, beta1, beta2, eps, weight_decay;
    float max_norm;  // <= 0: no clipping
};
struct AdamScalars {
    float coef, step_size, bc2_sqrt;
    int active;
};

// ---- optim.cu
int launch_sqnorm_stage1(const float* g, int64_t n, float* partial, BatchRef br, cudaStream_t st);
int launch_adam_prepare(const float* partial, int n_partial, const float* sqnorm_in, float* sqnorm_out,
                        AdamScalars* sc, AdamHyper hp, int64_t step_by_value, int* step_dev, const float* loss_rows,
                        const int32_t* n_targets_ptr, float* loss_out, BatchRef br, cudaStream_t st);
int launch_adam(float* w, float* g, float* m, float* v, int64_t n, const AdamScalars* sc, AdamHyper hp, bool zero_g,
                cudaStream_t st);

// ---- dense.cu   (act: 0 none, 1 tanh, 2 relu)
// Dropout source for the forward epilogue / backward mask: explicit keep bytes, or the counter-based generator.
struct Dropout {
    const uint8_t* keep = nullptr;       // 0/1 bytes, [rows x n]; nullptr -> counter-based generator
    const uint64_t* seed_dev = nullptr;  // generator seed (device)
    const int* step_dev = nullptr;       // device step counter mixed into the hash
    const int32_t* row_base = nullptr;   // engine mode: mask row of in-batch row m is row_base[b] + m
    int b = 0;
    float scale = 1.f;  // 1/(1-p)
    float p = 0.f;
    int enabled = 0;
};
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (uint32_t)x;
}
__device__ __forceinline__ float dropout_factor(const Dropout& d, int row, int col, int ld) {
    if (!d.enabled) return 1.f;
    int64_t r = (int64_t)row + (d.row_base ? d.row_base[d.b] : 0);
    if (d.keep != nullptr) return d.keep[r * ld + col] ? d.scale : 0.f;
    uint64_t seed = d.seed_dev ? *d.seed_dev : 0ull;
    uint64_t t = d.step_dev ? (uint64_t)(*d.step_dev) : 0ull;
    // keyed on (seed, optimizer step, row inside the batch, unit): independent of how batches are packed into plans
    uint32_t h = mix32(seed ^ (t * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)row << 24) ^ (uint64_t)col);
    return ((h >> 8) * (1.0f / 16777216.0f)) >= d.p ? d.scale : 0.f;
}

int launch_dense_fwd(const float* X, const float* W, const float* b, float* Y, float* Y_pre, Dropout drop, int m_max,
                     int n, int k, int act, BatchRef br, cudaStream_t st);
int launch_dense_bwd_x(const float* dY, const float* W, const float* A_prev, Dropout drop, float* dX, int m_max, int n,
                       int k, int act_prev, BatchRef br, cudaStream_t st);
int launch_dense_bwd_w(const float* dY, const float* X, float* dW, float* db, int m_max, int n, int k, BatchRef br,
                       cudaStream_t st);
int launch_colsum(const float* dY, int n, float* db, BatchRef br, cudaStream_t st);
// tcgen05 (UMMA, kind::tf32) variants; passes = 3: 3xTF32 (fp32-level accuracy), 1: single TF32 pass
int launch_dense_fwd_tc(const float* X, const float* W, const float* b, float* Y, float* Y_pre, Dropout drop, int m_max,
                        int n, int k, int act, int passes, BatchRef br, cudaStream_t st);
int launch_dense_bwd_x_tc(const float* dY, const float* W, const float* A_prev, Dropout drop, float* dX, int m_max,
                          int n, int k, int act_prev, int passes, BatchRef br, cudaStream_t st);
int launch_dense_bwd_w_tc(const float* dY, const float* X, float* dW, float* db, int m_max, int n, int k, int passes,
                          BatchRef br, cudaStream_t st);

// ---- ae.cu
int launch_ae_encoder_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* val,
                          const float* W1t, const float* b1, int H, float* A1, int n_rows_max, BatchRef br,
                          cudaStream_t st);
// ent_off: when non-null, per-batch-row offsets into the batch-ordered entry space (gout is written there);
// when null gout/pred are written at the CSR positions.
int launch_ae_decoder_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                          const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                          const int32_t* n_targets, const int32_t* ent_off, float* pred, float* gout, float* dZ3,
                          float* loss_rows, int tanh_deriv, int n_rows_max, BatchRef br, cudaStream_t st);

// ---- decoder_tc.cu: the decoder's last layer as tcgen05 GEMMs with CSR-scattered sparse operands (3xTF32)
// optional table of per-row CSR windows at 128-column tile borders (umma.cuh TileTab): mode 1 = indexed by the batch-row
// index, 2 = by the CSR row id; tab == null: the kernels binary-search instead
struct TcTab {
    const int32_t* tab;
    int mode;
};
int64_t decoder_tc_tab_ints(int64_t n_rows, int n_dec);
int build_tile_tab(const int32_t* rows, int n_rows, const int32_t* indptr, const int32_t* indices, int n_dec,
                   int32_t* tab, cudaStream_t st);
int decoder_tc_chunks_per_split(int n_rows_max, int n_dec, int H);
int decoder_tc_splits(int n_dec, int chunks_per_split);
int64_t decoder_tc_scratch_floats(int n_rows_max, int n_dec, int H);  // split-K partials followed by per-tile losses
int launch_decoder_tc_fwd(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                          const float* A3, const float* W4, const float* b4, int H, int n_dec, int loss_kind,
                          const int32_t* n_targets, const int32_t* ent_off, float* pred, float* gout, float* loss_part,
                          TcTab tab, int passes, int n_rows_max, BatchRef br, cudaStream_t st);
int launch_decoder_tc_bwd_a(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* gbuf,
                            const int32_t* ent_off, const float* A3, const float* W4, int H, int n_dec, float* part,
                            const float* loss_part, float* dZ3, float* loss_rows, int tanh_deriv, TcTab tab, int passes,
                            int n_rows_max, BatchRef br, cudaStream_t st);
int launch_decoder_tc_bwd_w(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* gbuf,
                            const int32_t* ent_off, const float* A3, int H, int n_dec, float* dW4, float* db4,
                            TcTab tab, int passes, int n_rows_max, BatchRef br, cudaStream_t st);

// ---- segments.cu
struct SegRef {  // segments [seg_lo, seg_hi) either by value or from device batch_seg_off[b], [b+1]
    const int32_t* batch_seg_off;
    const int32_t* n_seg_dev;
    int b;
    int64_t lo, hi;
    int32_t key_base_stride;  // column = seg_key - b * key_base_stride (0 for by-value use)
};
int launch_segment_reduce_rows(const int32_t* perm, const int32_t* seg_key, const int32_t* seg_off, SegRef sr,
                               int64_t n_seg_max, const float* coef, const int32_t* src_row, const float* src,
                               int width, float* grad, float* bias_grad, const int32_t* active, cudaStream_t st);
// Load-balanced segmented reduction of the engine: segments cut into chunks of <= kSegChunk entries.
constexpr int kSegChunk = 64;
struct ChunkedSegs {
    const int32_t* perm;
    const int32_t* ent_row;        // in-batch row of every entry (indexes src)
    const int32_t* seg_key;
    const int32_t* seg_off;
    const int32_t* batch_seg_off;  // [nb+1]
    const int32_t* seg_chunk_off;  // [n_seg+1] first chunk of every segment
    const int32_t* chunk_seg;      // [n_chunks] owning segment of every chunk
    const int32_t* batch_chunk_off;  // [nb+1] = seg_chunk_off[batch_seg_off[b]]
    float* part;                   // [max chunks per batch x width] partial rows of multi-chunk segments
    float* part_bias;
    int b;
    int n_cols;
};
int launch_segment_chunks(ChunkedSegs cs, int n_chunk_max, int n_seg_max, const float* coef, const float* src,
                          int width, float* grad, float* bias_grad, const int32_t* active, cudaStream_t st);
int build_seg_chunks(const int32_t* seg_off, const int32_t* n_seg, int64_t cap, int32_t* n_ch, int32_t* seg_chunk_off,
                     int32_t* chunk_seg, void* temp, int64_t temp_bytes, cudaStream_t st);

// Decoder of the engine: batch rows cut into chunks of <= kDecChunk target entries (heavy rows span several blocks).
// 128 is measured: 64 -> 264 ms per ML1M round (more prologues and partial rows), 256 -> 220 ms, 128 -> 215 ms.
constexpr int kDecChunk = 128;
struct DecChunks {
    const int32_t* chunk_off;        // [rows+1] first chunk of every batch-row (epoch-wide numbering)
    const int32_t* chunk_row;        // [n_chunks] batch-row (epoch-wide index) of every chunk
    const int32_t* batch_chunk_off;  // [nb+1]
    float* dz_part;                  // [max chunks per batch x H]
    float* loss_part;                // [max chunks per batch]
};
int launch_ae_decoder_chunks(const int32_t* rows, const int32_t* indptr, const int32_t* indices, const float* target,
                             const float* A3, const float* W4, const float* b4, int H, int loss_kind,
                             const int32_t* n_targets, const int32_t* ent_off, DecChunks dc, float* gout, float* dZ3,
                             float* loss_rows, int n_rows_max, BatchRef br, cudaStream_t st, int blocks_hint = 0);

// ---- fused.cu: the six-launch training step (see the header comment of fused.cu)
constexpr int kFusedRows = 8;        // batch rows per CTA of the row-local forward / backward kernels
constexpr int kDbPartStride = 768;   // per row-CTA column-sum partials: db3 [256] | db2 halves [2 x 128] | db1 [256]
constexpr int kDwSlice = 64;         // batch rows per split of the dW3 / dW2 tiles (<= 8 slices: batch_rows <= 512)
constexpr int kFusedMaxRowChunks = 4095;  // 12-bit chunk counters in the decoder chunk metadata

struct FusedFwd {
    BatchRef br;
    const int32_t *rows, *d_indptr, *d_indices;
    const float* d_val;
    const float *W1t, *b1, *W2t, *b2, *W3t, *b3;
    float *a1, *a2, *c, *a3;  // a1 / a2 / c may be null (prediction)
    Dropout drop;
};
struct FusedDec {
    BatchRef br;
    const int4* meta;                // per chunk, plan_dec_meta_kernel
    const int32_t* batch_chunk_off;  // [nb + 1]
    const int32_t* n_targets;        // [nb] targets per batch
    const int32_t* t_indices;
    const float* target;
    const int32_t* inv_perm;         // batch-order entry id -> position in the (batch, column)-sorted order
    const float *A3, *W4, *b4;
    float *g_sorted, *dZ3, *loss_rows, *dz_part, *loss_part;
    int* row_cnt;                    // [batch_rows] arrival counters, zero between steps
};
struct FusedBwd {
    BatchRef br;
    const int32_t* t_len;            // targets per batch-row (epoch-wide index)
    const float *dz3, *W3, *W2, *a1, *a2;
    float *dz2, *dz1, *part_db;
    Dropout drop;
};
struct FusedSeg {
    const int4* meta;                // per chunk: {e0, e1, output row, k | n_chunks << 16}
    const int32_t* batch_chunk_off;
    const int32_t* row_sorted;       // in-batch source row of every sorted entry
    const float* coef_sorted;        // coefficient of every sorted entry
    float* part;
    float* part_bias;
    int* cnt;                        // arrival counters at the first partial slot of a segment, zero between steps
    const int32_t* active;
    int b;
};
struct FusedGrad {
    BatchRef br;
    const int32_t* t_len;
    const float *dz3, *dz2, *c, *a1;
    const float* part_db;
    float* G;
    int64_t oW2, oW3, ob1, ob2, ob3;
    float* part_dw;                  // [16 tiles x 8 slices x 64 x 64]
    int* dw_cnt;                     // [16]
    int rows_per_cta;                // row tile of the backward-rows kernel that wrote part_db
};
struct FusedPlanSide {
    int64_t n_entries;
    int n_cols;
    const int32_t *n_seg, *seg_off, *seg_key, *seg_chunk_off, *perm, *ent_row;
    const float* val_ord;            // null on the target side
    int4* seg_meta;
    int32_t* row_sorted;
    float* val_sorted;               // null on the target side
    int32_t* inv_perm;               // null on the data side
};
struct FusedPlanArgs {
    int64_t dec_chunk_cap;
    int n_rows;
    const int32_t *t_chunk_off, *t_chunk_row, *rows, *row_off, *row_batch, *t_indptr, *t_ent_off;
    int4* dec_meta;
    FusedPlanSide t, d;
};
int launch_fused_fwd(const FusedFwd& p, int n_rows_max, cudaStream_t st, bool pdl = false, int R = kFusedRows);
// gather: 0 = rows through registers (plain loads), 1 = rows through shared-memory rings (bulk copies, bulk.cuh)
int launch_fused_dec(const FusedDec& p, int blocks_hint, int gather, cudaStream_t st, bool pdl = false);
// fused_rows.cu: the row kernels with W2 / W3 streamed through shared memory by 32 KB bulk copies; R = rows per CTA
int fused_rows_per_cta(int batch_rows);
int prepare_fused_rows();
int launch_fused_fwd_tma(const FusedFwd& p, int n_rows_max, int R, cudaStream_t st, bool pdl = false);
int launch_fused_bwd_rows_tma(const FusedBwd& p, int n_rows_max, int R, cudaStream_t st, bool pdl = false);
int launch_fused_bwd_rows(const FusedBwd& p, int n_rows_max, cudaStream_t st, bool pdl = false, int R = kFusedRows);
int launch_fused_seg_chunks(const FusedSeg& s, const float* src, float* grad, float* bias_grad, int n_chunk_max,
                            int gather, cudaStream_t st);
int launch_fused_grad_phase(const FusedGrad& p, const FusedSeg& s, const float* src, float* grad, int n_chunk_max,
                            cudaStream_t st, bool pdl = false);
int launch_norm_prepare(const float* g, int64_t n, float* partial, AdamScalars* sc, int* step_dev,
                        const float* loss_rows, const int32_t* t_len, const int32_t* n_targets_ptr, float* loss_out,
                        BatchRef br, cudaStream_t st, bool pdl = false);
int launch_adam_shadow(float* w, float* g, float* m, float* v, int64_t n, const AdamScalars* sc, AdamHyper hp,
                       const float* partial, const int* step_dev, int64_t oW2, int64_t oW3, float* W2t, float* W3t,
                       cudaStream_t st, bool pdl = false);
int launch_shadow_refresh(const float* W2, const float* W3, float* W2t, float* W3t, cudaStream_t st);
int launch_plan_fused(const FusedPlanArgs& a, cudaStream_t st);

// ---------------------------------------------------------------- organization groups (one launch = all organizations)
// Device-visible view of one organization. A group launch adds the organization as grid dimension z, so a step of
// ALL organizations of a rank is the same ~20 launches as a step of one (the per-organization kernels are far too
// small to fill 148 SMs, and ~2.7 us of launch processing per kernel node is what bounded the per-org design).
struct OrgDev {
    const int32_t *rows, *row_off, *active;
    const int32_t *d_indptr, *d_indices, *t_indptr, *t_indices;
    const float *d_val, *t_val;
    float *P, *G, *M, *V;
    int64_t n_params, oW1, ob1, oW2, ob2, oW3, ob3, oW4, ob4;
    int n_enc, n_dec;
    float *a1, *a2, *c, *a3, *dz3, *dz2, *dz1, *loss_rows;
    const int32_t *t_ent_off, *t_batch_cnt;
    DecChunks dc;
    float* gbuf;
    ChunkedSegs seg_t, seg_d;  // .b is overwritten per launch
    const float* dval_ord;
    float* partial;
    AdamScalars* sc;
    int* step_dev;
    float* loss_buf;
    const uint64_t* seed_dev;
};
int launch_group_encoder(const OrgDev* orgs, int G, int b, int B, int H1, cudaStream_t st);
int launch_group_dense(const OrgDev* orgs, int G, int b, int B, int H1, int H2, int which, cudaStream_t st);
int launch_group_decoder(const OrgDev* orgs, int G, int b, int B, int H1, cudaStream_t st);
int launch_group_segments(const OrgDev* orgs, int G, int b, int side, int n_cols_max, int H1, cudaStream_t st);
int launch_group_colsum_dz1(const OrgDev* orgs, int G, int b, int H1, cudaStream_t st);
int launch_group_optim(const OrgDev* orgs, int G, int b, int64_t n_params_max, AdamHyper hp, cudaStream_t st);

int64_t sort_segments_temp_bytes(int64_t n);
int sort_segments(const uint32_t* keys, int64_t n, int key_bits, int32_t* perm, int32_t* seg_key, int32_t* seg_off,
                  int32_t* n_seg, void* temp, int64_t temp_bytes, cudaStream_t st);

}  // namespace dmt
