// Small dense fp32 layers of the AAE (256 <-> 128 wide; reference src/models/ae.py:9-59, called at :122,:133) with
// bias / tanh / dropout fused into the GEMM epilogue and the activation derivative fused into the backward GEMM.
// fp32 FFMA on purpose: the parity bar is 1e-5 relative on the loss (BASELINE.json north_star), which a single-pass
// bf16/TF32 tensor-core product cannot hold; these GEMMs are <10 % of a step.
#include "kernels.cuh"
#include "umma.cuh"

namespace dmt {

constexpr int BK = 32;

// C = op(A) op(B) with a fused epilogue. 256 threads, BM x BN tile, (BM/16) x (BN/16) outputs per thread.
// Global -> register prefetch of k-tile t+1 overlaps the FFMA loop on k-tile t (double-buffered shared memory, one
// barrier per k-tile): these GEMMs are tiny (M = one 500-row batch), so latency, not bandwidth, is what is hidden.
template <int BM, int BN, bool A_KMAJOR, bool B_KMAJOR, int DYN /*0: M dynamic, 1: K dynamic*/, class Epi>
__device__ __forceinline__ void sgemm_body(const float* __restrict__ A, const float* __restrict__ B, int M,
                                                    int N, int K, int lda, int ldb, Epi epi, BatchRef br, int zK) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int LA = BM * BK / 256, LB = BN * BK / 256;
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    if (DYN == 0) M = hi - lo; else K = hi - lo;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[LA], rb[LB];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (A_KMAJOR) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = A_KMAJOR ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int idx = tid + i * 256;
            int n, k;
            if (B_KMAJOR) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < K) v = B_KMAJOR ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
            rb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (A_KMAJOR) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int idx = tid + i * 256;
            int n, k;
            if (B_KMAJOR) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            Bs[buf][k][n] = rb[i];
        }
    };
    // split-K (grid.z slices of zK along K, partial results at epi's z offset) for reductions over many rows
    int k_base = 0;
    if (zK > 0) {
        k_base = blockIdx.z * zK;
        K = min(K, k_base + zK);
    }
    const int nk = (K - k_base + BK - 1) / BK;
    if (nk > 0) {
        gload(k_base);
        sstore(0);
    }
    __syncthreads();
    for (int t = 0; t < nk; ++t) {
        const int buf = t & 1;
        if (t + 1 < nk) gload(k_base + (t + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (t + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * TN + j;
            if (gn < N) epi(gm, gn, acc[i][j]);
        }
    }
}

__device__ __forceinline__ float act_apply(int act, float v) {
    if (act == 1) return tanhf(v);
    if (act == 2) return fmaxf(v, 0.f);
    return v;
}
__device__ __forceinline__ float act_deriv(int act, float a) {
    if (act == 1) return 1.f - a * a;
    if (act == 2) return a > 0.f ? 1.f : 0.f;
    return 1.f;
}

struct FwdEpi {
    const float* b;
    float* Y;
    float* Ypre;
    Dropout drop;
    int ld, act;
    __device__ __forceinline__ void operator()(int m, int n, float acc) const {
        float v = act_apply(act, acc + (b ? b[n] : 0.f));
        if (drop.enabled) {
            if (Ypre) Ypre[(int64_t)m * ld + n] = v;
            v *= dropout_factor(drop, m, n, ld);
        }
        Y[(int64_t)m * ld + n] = v;
    }
};
struct BwdXEpi {
    const float* A_prev;
    float* dX;
    Dropout drop;
    int ld, act;
    __device__ __forceinline__ void operator()(int m, int n, float acc) const {
        float d = acc;
        if (drop.enabled) d *= dropout_factor(drop, m, n, ld);
        if (A_prev) d *= act_deriv(act, A_prev[(int64_t)m * ld + n]);
        dX[(int64_t)m * ld + n] = d;
    }
};
struct StoreEpi {
    float* C;
    int ld;
    int64_t zstride;  // split-K: slice z writes its partial at C + z * zstride
    __device__ __forceinline__ void operator()(int m, int n, float acc) const {
        C[(int64_t)blockIdx.z * zstride + (int64_t)m * ld + n] = acc;
    }
};

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t count,
                                                            float* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[(int64_t)z * count + i];  // fixed order: deterministic
        out[i] = s;
    }
}

// db[n] = sum over the batch rows of dY[:, n]. 32 columns x 32 row-lanes per block: the row loop is 32x shorter than
// the batch (latency, not bandwidth, is what this tiny reduction costs).
__device__ __forceinline__ void colsum_body(const float* __restrict__ dY, int n, float* __restrict__ db,
                                                      BatchRef br) {
    __shared__ float sh[32][33];
    int lo, hi;
    if (!batch_range(br, lo, hi)) return;
    int m = hi - lo;
    int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    int col = blockIdx.x * 32 + x;
    float s = 0.f;
    if (col < n)
        for (int r = y; r < m; r += 32) s += dY[(int64_t)r * n + col];
    sh[y][x] = s;
    __syncthreads();
    if (y == 0 && col < n) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += sh[i][x];
        db[col] = t;
    }
}

template <int BM, int BN, bool A_KMAJOR, bool B_KMAJOR, int DYN, class Epi>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* A, const float* B, int M, int N, int K, int lda,
                                                    int ldb, Epi epi, BatchRef br, int zK) {
    sgemm_body<BM, BN, A_KMAJOR, B_KMAJOR, DYN, Epi>(A, B, M, N, K, lda, ldb, epi, br, zK);
}
__global__ void __launch_bounds__(1024) colsum_kernel(const float* dY, int n, float* db, BatchRef br) {
    colsum_body(dY, n, db, br);
}

// Group forms (z = organization). `which`: 0 a2/c = tanh(a1 W2^T + b2) with dropout, 1 a3 = tanh(c W3^T + b3),
// 2 dW3 = dz3^T c, 3 dz2 = (dz3 W3) * drop * (1-a2^2), 4 dW2 = dz2^T a1, 5 dz1 = (dz2 W2) * (1-a1^2)
__device__ __forceinline__ Dropout group_dropout(const OrgDev& o, int b) {
    Dropout d;
    d.seed_dev = o.seed_dev;
    d.step_dev = o.step_dev;
    d.row_base = o.row_off;
    d.b = b;
    d.scale = 2.0f;  // nn.Dropout(p=0.5), reference src/models/ae.py:81
    d.p = 0.5f;
    d.enabled = 1;
    return d;
}
template <int WHICH, int BM, int BN>
__global__ void __launch_bounds__(256) dense_group(const OrgDev* __restrict__ orgs, int b, int B, int H1, int H2) {
    const OrgDev& o = orgs[blockIdx.z];
    const BatchRef br{o.row_off, o.active, b, 0, 0};
    float *W2 = o.P + o.oW2, *b2 = o.P + o.ob2, *W3 = o.P + o.oW3, *b3 = o.P + o.ob3;
    if (WHICH == 0) {
        FwdEpi epi{b2, o.c, o.a2, group_dropout(o, b), H2, 1};
        sgemm_body<BM, BN, true, true, 0>(o.a1, W2, B, H2, H1, H1, H1, epi, br, 0);
    } else if (WHICH == 1) {
        FwdEpi epi{b3, o.a3, nullptr, Dropout(), H1, 1};
        sgemm_body<BM, BN, true, true, 0>(o.c, W3, B, H1, H2, H2, H2, epi, br, 0);
    } else if (WHICH == 2) {
        StoreEpi epi{o.G + o.oW3, H2, 0};
        sgemm_body<BM, BN, false, false, 1>(o.dz3, o.c, H1, H2, B, H1, H2, epi, br, 0);
    } else if (WHICH == 3) {
        BwdXEpi epi{o.a2, o.dz2, group_dropout(o, b), H2, 1};
        sgemm_body<BM, BN, true, false, 0>(o.dz3, W3, B, H2, H1, H1, H2, epi, br, 0);
    } else if (WHICH == 4) {
        StoreEpi epi{o.G + o.oW2, H1, 0};
        sgemm_body<BM, BN, false, false, 1>(o.dz2, o.a1, H2, H1, B, H2, H1, epi, br, 0);
    } else {
        BwdXEpi epi{o.a1, o.dz1, Dropout(), H1, 1};
        sgemm_body<BM, BN, true, false, 0>(o.dz2, W2, B, H1, H2, H2, H1, epi, br, 0);
    }
}
// column sums: which 2 -> db3 from dz3 [H1], 4 -> db2 from dz2 [H2], 6 -> db1 from dz1 [H1]
__global__ void __launch_bounds__(1024) colsum_group(const OrgDev* __restrict__ orgs, int b, int which, int n) {
    const OrgDev& o = orgs[blockIdx.z];
    const float* src = which == 2 ? o.dz3 : (which == 4 ? o.dz2 : o.dz1);
    float* dst = o.G + (which == 2 ? o.ob3 : (which == 4 ? o.ob2 : o.ob1));
    colsum_body(src, n, dst, BatchRef{o.row_off, o.active, b, 0, 0});
}

int launch_group_dense(const OrgDev* orgs, int G, int b, int B, int H1, int H2, int which, cudaStream_t st) {
    // output shapes: 0 [B x H2], 1 [B x H1], 2 [H1 x H2], 3 [B x H2], 4 [H2 x H1], 5 [B x H1]
    const int M = (which == 2) ? H1 : (which == 4 ? H2 : B);
    const int N = (which == 0 || which == 3 || which == 2) ? H2 : H1;
    dim3 grid((N + 63) / 64, (M + 63) / 64, G);  // 64x64 tiles: with G organizations per launch the grid is wide enough
#define DMT_DG(W) dense_group<W, 64, 64><<<grid, 256, 0, st>>>(orgs, b, B, H1, H2)
    switch (which) {
        case 0: DMT_DG(0); break;
        case 1: DMT_DG(1); break;
        case 2: DMT_DG(2); break;
        case 3: DMT_DG(3); break;
        case 4: DMT_DG(4); break;
        default: DMT_DG(5); break;
    }
#undef DMT_DG
    DMT_LAUNCH_CHECK();
    if (which == 2 || which == 4) {
        int n = which == 2 ? H1 : H2;
        colsum_group<<<dim3((n + 31) / 32, 1, G), 1024, 0, st>>>(orgs, b, which, n);
        DMT_LAUNCH_CHECK();
    }
    return 0;
}

int launch_group_colsum_dz1(const OrgDev* orgs, int G, int b, int H1, cudaStream_t st) {
    colsum_group<<<dim3((H1 + 31) / 32, 1, G), 1024, 0, st>>>(orgs, b, 6, H1);
    DMT_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- tcgen05 (UMMA) path, 3xTF32 for fp32 parity
// C[M x N] = A . B^T on the 5th-generation tensor cores (building blocks and the operand layout: umma.cuh).
// A(m, k) = A[m * lda + k] when A_KCONTIG else A[k * lda + m]; B(n, k) likewise.
template <bool A_KCONTIG, bool B_KCONTIG, int DYN, class Epi>
__global__ void __launch_bounds__(umma::kThreads, 1) umma_gemm_kernel(const float* __restrict__ A, int64_t lda,
                                                                      const float* __restrict__ B, int64_t ldb, int M,
                                                                      int N, int K, Epi epi, BatchRef br, int passes) {
    extern __shared__ uint8_t umma_smem_raw[];
    int lo_, hi_;
    if (!batch_range(br, lo_, hi_)) return;
    if (DYN == 0) M = hi_ - lo_; else if (DYN == 1) K = hi_ - lo_;
    const int m0 = blockIdx.y * umma::TM, n0 = blockIdx.x * umma::TN;
    if (m0 >= M || n0 >= N) return;  // uniform per CTA, before any allocation or barrier
    const umma::Pipe pp = umma::pipe_setup(umma_smem_raw, 0, (K + umma::TK - 1) / umma::TK);
    const int warp = threadIdx.x >> 5;
    if (warp < 8) {
        const int team = warp >> 2, tt = threadIdx.x & (umma::kTeam - 1);
        const int n_mine = team == 0 ? pp.n0 : pp.n1, first = team == 0 ? 0 : pp.n0;
        for (int i = 0; i < n_mine; ++i) {
            const int c = umma::chunk_slot(pp, team, i), k0 = (first + i) * umma::TK;
            const umma::Stage st = umma::producer_acquire(pp, c);
            if (A_KCONTIG) umma::stage_kcontig(tt, A, lda, m0, M, k0, K, st.A_hi, st.A_lo, passes);
            else umma::stage_transposed(tt, A, lda, m0, M, k0, K, st.A_hi, st.A_lo, passes);
            if (B_KCONTIG) umma::stage_kcontig(tt, B, ldb, n0, N, k0, K, st.B_hi, st.B_lo, passes);
            else umma::stage_transposed(tt, B, ldb, n0, N, k0, K, st.B_hi, st.B_lo, passes);
            umma::producer_commit(pp, c);
        }
        umma::wait_accumulator(pp);
        // epilogue: accumulator row = TMEM lane; warps w and w + 4 split the 128 columns
        const int row = m0 + umma::epi_row(), c0 = umma::epi_col0();
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 32) {
            float v[32];
            if (K > 0) {
                umma::load_acc32(pp, c0 + cc, v);  // warp-collective: every lane takes part
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (row < M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c0 + cc + j;
                    if (col < N) epi(row, col, v[j]);
                }
            }
        }
    } else if (warp == umma::kMmaWarp) {
        umma::mma_loop(pp, passes);
    }
    umma::pipe_teardown(pp);
}

template <bool AK, bool BKC, int DYN, class Epi>
static int launch_umma(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K, Epi epi,
                       BatchRef br, int passes, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    auto kern = umma_gemm_kernel<AK, BKC, DYN, Epi>;
    static bool configured = false;  // per instantiation
    if (!configured) {
        DMT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, umma::smem_bytes(0)));
        configured = true;
    }
    dim3 grid((N + umma::TN - 1) / umma::TN, (M + umma::TM - 1) / umma::TM);
    kern<<<grid, umma::kThreads, umma::smem_bytes(0), st>>>(A, lda, B, ldb, M, N, K, epi, br, passes);
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_dense_fwd_tc(const float* X, const float* W, const float* b, float* Y, float* Y_pre, Dropout drop, int m_max,
                        int n, int k, int act, int passes, BatchRef br, cudaStream_t st) {
    FwdEpi epi{b, Y, Y_pre, drop, n, act};
    return launch_umma<true, true, 0>(X, k, W, k, m_max, n, k, epi, br, passes, st);
}

int launch_dense_bwd_x_tc(const float* dY, const float* W, const float* A_prev, Dropout drop, float* dX, int m_max,
                          int n, int k, int act_prev, int passes, BatchRef br, cudaStream_t st) {
    BwdXEpi epi{A_prev, dX, drop, k, act_prev};
    // dX[m x k] = dY[m x n] . W[n x k]: B(col = k index, reduction = n index) = W[n_idx * k + k_idx]
    return launch_umma<true, false, 0>(dY, n, W, k, m_max, k, n, epi, br, passes, st);
}

int launch_dense_bwd_w_tc(const float* dY, const float* X, float* dW, float* db, int m_max, int n, int k, int passes,
                          BatchRef br, cudaStream_t st) {
    StoreEpi epi{dW, k, 0};
    // dW[n x k] = dY^T X: A(row = n index, reduction = batch row) = dY[r * n + n_idx]; B(col = k index, r) = X[r * k + k_idx]
    int rc = launch_umma<false, false, 1>(dY, n, X, k, n, k, m_max, epi, br, passes, st);
    if (rc) return rc;
    if (db != nullptr) return launch_colsum(dY, n, db, br, st);
    return 0;
}

template <bool AK, bool BKM, int DYN, class Epi>
static int launch_gemm(const float* A, const float* B, int M, int N, int K, int lda, int ldb, Epi epi, BatchRef br,
                       cudaStream_t st, int splits = 1, int zK = 0) {
    if (M <= 0 || N <= 0) return 0;
    // Small problems (one 500-row batch) take 32x32 tiles: measured on B200 they finish a layer in ~9 us against ~17 us
    // for 64x64 tiles (4x fewer blocks, each 4x longer) — per-step latency is what bounds an organization's epoch.
    int64_t big = (int64_t)((M + 63) / 64) * ((N + 63) / 64) * splits;
    if (big >= kNumSMs) {
        dim3 grid((N + 63) / 64, (M + 63) / 64, splits);
        sgemm_kernel<64, 64, AK, BKM, DYN, Epi><<<grid, 256, 0, st>>>(A, B, M, N, K, lda, ldb, epi, br, zK);
    } else {
        dim3 grid((N + 31) / 32, (M + 31) / 32, splits);
        sgemm_kernel<32, 32, AK, BKM, DYN, Epi><<<grid, 256, 0, st>>>(A, B, M, N, K, lda, ldb, epi, br, zK);
    }
    DMT_LAUNCH_CHECK();
    return 0;
}

int launch_dense_fwd(const float* X, const float* W, const float* b, float* Y, float* Y_pre, Dropout drop, int m_max,
                     int n, int k, int act, BatchRef br, cudaStream_t st) {
    FwdEpi epi{b, Y, Y_pre, drop, n, act};
    return launch_gemm<true, true, 0>(X, W, m_max, n, k, k, k, epi, br, st);
}

int launch_dense_bwd_x(const float* dY, const float* W, const float* A_prev, Dropout drop, float* dX, int m_max, int n,
                       int k, int act_prev, BatchRef br, cudaStream_t st) {
    BwdXEpi epi{A_prev, dX, drop, k, act_prev};
    return launch_gemm<true, false, 0>(dY, W, m_max, k, n, n, k, epi, br, st);
}

int launch_dense_bwd_w(const float* dY, const float* X, float* dW, float* db, int m_max, int n, int k, BatchRef br,
                       cudaStream_t st) {
    int rc;
    if (br.row_off == nullptr && m_max > 4096) {
        // many rows (NCF: one row per rating): split the reduction over grid.z, then add the partials in order
        int splits = (m_max + 2047) / 2048;
        if (splits > 64) splits = 64;
        int zK = (m_max + splits - 1) / splits;
        zK = (zK + BK - 1) / BK * BK;
        splits = (m_max + zK - 1) / zK;
        float* part = nullptr;
        int64_t count = (int64_t)n * k;
        DMT_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), (size_t)splits * count * sizeof(float), st));
        StoreEpi pepi{part, k, count};
        rc = launch_gemm<false, false, 1>(dY, X, n, k, m_max, n, k, pepi, br, st, splits, zK);
        if (rc == 0) {
            int blocks = (int)((count + 255) / 256);
            if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
            splitk_reduce_kernel<<<blocks, 256, 0, st>>>(part, splits, count, dW);
            count_launch(1);
        }
        cudaFreeAsync(part, st);
        if (rc) return rc;
        DMT_CUDA(cudaGetLastError());
    } else {
        StoreEpi epi{dW, k, 0};
        rc = launch_gemm<false, false, 1>(dY, X, n, k, m_max, n, k, epi, br, st);
        if (rc) return rc;
    }
    if (db != nullptr) return launch_colsum(dY, n, db, br, st);
    return 0;
}

int launch_colsum(const float* dY, int n, float* db, BatchRef br, cudaStream_t st) {
    colsum_kernel<<<(n + 31) / 32, 1024, 0, st>>>(dY, n, db, br);
    DMT_LAUNCH_CHECK();
    return 0;
}

}  // namespace dmt

using namespace dmt;

extern "C" {

int dmt_dense_fwd(const float* X, const float* W, const float* b, float* Y, float* Y_pre, const uint8_t* keep,
                  float keep_scale, int m, int n, int k, int act, void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0 && act >= 0 && act <= 2, "dmt_dense_fwd: bad argument");
    Dropout d;
    d.keep = keep;
    d.scale = keep_scale;
    d.enabled = keep != nullptr;
    return launch_dense_fwd(X, W, b, Y, Y_pre, d, m, n, k, act, batch_by_value(0, m), as_stream(stream));
}

int dmt_dense_bwd_x(const float* dY, const float* W, const float* A_prev, const uint8_t* keep, float keep_scale,
                    float* dX, int m, int n, int k, int act_prev, void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0 && act_prev >= 0 && act_prev <= 2, "dmt_dense_bwd_x: bad argument");
    Dropout d;
    d.keep = keep;
    d.scale = keep_scale;
    d.enabled = keep != nullptr;
    return launch_dense_bwd_x(dY, W, A_prev, d, dX, m, n, k, act_prev, batch_by_value(0, m), as_stream(stream));
}

int dmt_dense_fwd_tc(const float* X, const float* W, const float* b, float* Y, float* Y_pre, const uint8_t* keep,
                     float keep_scale, int m, int n, int k, int act, int passes, void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0 && act >= 0 && act <= 2 && (passes == 1 || passes == 3),
                "dmt_dense_fwd_tc: bad argument");
    Dropout d;
    d.keep = keep;
    d.scale = keep_scale;
    d.enabled = keep != nullptr;
    return launch_dense_fwd_tc(X, W, b, Y, Y_pre, d, m, n, k, act, passes, batch_by_value(0, m), as_stream(stream));
}

int dmt_dense_bwd_x_tc(const float* dY, const float* W, const float* A_prev, const uint8_t* keep, float keep_scale,
                       float* dX, int m, int n, int k, int act_prev, int passes, void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0 && (passes == 1 || passes == 3), "dmt_dense_bwd_x_tc: bad argument");
    Dropout d;
    d.keep = keep;
    d.scale = keep_scale;
    d.enabled = keep != nullptr;
    return launch_dense_bwd_x_tc(dY, W, A_prev, d, dX, m, n, k, act_prev, passes, batch_by_value(0, m),
                                 as_stream(stream));
}

int dmt_dense_bwd_w_tc(const float* dY, const float* X, float* dW, float* db, int m, int n, int k, int passes,
                       void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0 && (passes == 1 || passes == 3), "dmt_dense_bwd_w_tc: bad argument");
    return launch_dense_bwd_w_tc(dY, X, dW, db, m, n, k, passes, batch_by_value(0, m), as_stream(stream));
}

int dmt_dense_bwd_w(const float* dY, const float* X, float* dW, float* db, int m, int n, int k, void* stream) {
    DMT_REQUIRE(m >= 0 && n > 0 && k > 0, "dmt_dense_bwd_w: bad argument");
    return launch_dense_bwd_w(dY, X, dW, db, m, n, k, batch_by_value(0, m), as_stream(stream));
}

}  // extern "C"
