// tcgen05 (UMMA) building blocks for sm_100a: fp32 GEMM tiles on the 5th-generation tensor cores with 3xTF32.
//
// D[128 x 128] (fp32, in TMEM) += A[128 x 32] . B[128 x 32]^T per k-chunk, issued by ONE thread as
// tcgen05.mma.cta_group::1.kind::tf32 (4 k-steps of 8 per chunk). fp32 parity needs more than one TF32 pass
// (10-bit mantissa): every operand element x is split while it is staged into shared memory into
//   hi = rna_tf32(x),  lo = x - hi   (exact in fp32)
// and three MMAs (lo*hi, hi*lo, hi*hi) accumulate into the same TMEM tile ("3xTF32", ~2^-21 relative error).
// passes = 1 skips the lo terms (a labelled reduced-precision mode).
//
// Operands are staged by the CTA itself — they have to be transformed (split, transposed, or scattered from a CSR)
// anyway, so TMA cannot deliver them — into the canonical K-major SWIZZLE_128B layout that the UMMA shared-memory
// descriptor describes (CUTLASS cute/atom/mma_traits_sm100.hpp, "Swizzle<3,4,3> o ((8,n),2):((8,SBO),1)" in 16-byte
// units): a tile is 128 rows x 128 B (32 fp32 along K); row r lives at (r/8)*1024 + (r%8)*128, its 16-byte chunk j is
// stored at chunk j ^ (r%8); one tf32 MMA consumes K = 8 elements = 32 B, so the descriptor start address advances by
// 32 B per k-step inside the 128 B swizzle atom.
//
// CTA = 10 warps, warp-specialised:
//   warps 0-3 / 4-7  two producer TEAMS of 128 threads; team 0 stages the first half of the CTA's k-chunks, team 1 the
//                    second half, interleaved in the consumption order 0,1,0,1,...  A team has two MMA-chunk times
//                    to cover its global-load latency. Generic-proxy stores -> fence.proxy.async -> mbarrier arrive.
//   warp 8           lane 0 waits for a full stage, issues its 12 (4) MMAs and commits them to the stage's "empty"
//                    mbarrier; a last commit signals "accumulator complete". Also owns the TMEM allocation.
//   warp 9           auxiliary (per-row CSR windows for the epilogue), overlapped with the main loop.
// A ring of 3 stages x 64 KB (A_hi, A_lo, B_hi, B_lo) keeps the tensor pipe fed; the epilogue reads the accumulator
// with tcgen05.ld (warp w and w+4 share TMEM lanes 32*(w%4).. and split the 128 columns).
#pragma once
#include "common.cuh"

namespace dmt {
namespace umma {

constexpr int TM = 128, TN = 128, TK = 32;
constexpr int kStages = 3;
constexpr int kTeam = 128;                     // threads of one producer team
constexpr int kProducers = 2 * kTeam;          // warps 0-7
constexpr int kMmaWarp = 8, kAuxWarp = 9;
constexpr int kThreads = 320;
constexpr int kTileBytes = TM * 128;           // one operand, one of hi/lo, one k-chunk: 16 KB
constexpr int kStageBytes = 4 * kTileBytes;    // A_hi, A_lo, B_hi, B_lo
constexpr int kRingBytes = kStages * kStageBytes;
constexpr int kCtrlBytes = 128;                // mbarriers + TMEM address slot
// dynamic shared memory of a kernel that keeps `extra` bytes (16 B multiple) of its own behind the ring
constexpr int smem_bytes(int extra) { return 1024 /*alignment slack*/ + kRingBytes + extra + kCtrlBytes; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major),
// SBO>>4 [32,46) = 1024 B between 8-row groups, version = 1 [46,48), base_offset = 0 (tiles are 1024 B aligned),
// layout_type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// InstrDescriptor: c_format F32 = 1 [4,6), a/b_format TF32 = 2 [7,10)/[10,13), both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    // arrives on the mbarrier once every MMA issued so far by this thread has finished reading smem / writing TMEM
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded: an arrival that never comes (a malformed descriptor, a lost MMA) traps after ~seconds instead of hanging
// the GPU; the host then sees a launch failure.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void team_sync(int team) {  // named barrier of one producer team (ids 1, 2)
    if (team == 0) asm volatile("bar.sync 1, %0;" ::"r"(kTeam) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(kTeam) : "memory");
}
__device__ __forceinline__ void producers_sync() {  // all 8 producer warps
    asm volatile("bar.sync 3, %0;" ::"r"(kProducers) : "memory");
}

// byte offset of element (r, k) of a [128 x 32] fp32 tile in the swizzled layout
__device__ __forceinline__ uint32_t tile_off(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7))) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ uint32_t tf32_hi(float v) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    return h;
}

struct Stage {
    uint8_t *A_hi, *A_lo, *B_hi, *B_lo;
};

struct Pipe {
    uint8_t* base;   // 1024 B aligned start of the ring (the epilogue may reuse it once the accumulator is complete)
    uint8_t* extra;  // the kernel's own area behind the ring
    uint32_t full0, empty0, done, tmem;  // shared-memory addresses of the mbarriers (stage s: +8*s), TMEM base
    int n0, n1;      // chunks of team 0 / team 1 (n1 <= n0 <= n1 + 1)
};

__device__ __forceinline__ Stage stage_of(const Pipe& p, int c) {
    uint8_t* s = p.base + (c % kStages) * kStageBytes;
    return Stage{s, s + kTileBytes, s + 2 * kTileBytes, s + 3 * kTileBytes};
}

// All 320 threads, after every early exit. n_chunks = k-chunks this CTA accumulates.
__device__ __forceinline__ Pipe pipe_setup(uint8_t* raw, int extra, int n_chunks) {
    Pipe p;
    p.base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    p.extra = p.base + kRingBytes;
    uint8_t* ctrl = p.extra + extra;
    p.full0 = smem_u32(ctrl);
    p.empty0 = p.full0 + 8 * kStages;
    p.done = p.empty0 + 8 * kStages;
    uint32_t* slot = reinterpret_cast<uint32_t*>(ctrl + 8 * (2 * kStages + 1));
    p.n0 = (n_chunks + 1) >> 1;
    p.n1 = n_chunks >> 1;
    if ((threadIdx.x >> 5) == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
                     "r"((uint32_t)TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(p.full0 + 8 * s, kTeam);
            mbar_init(p.empty0 + 8 * s, 1);
        }
        mbar_init(p.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    p.tmem = *slot;
    return p;
}

// consumption index of a team's i-th chunk: 0,1,0,1,... while both teams have chunks, then team 0's remainder
__device__ __forceinline__ int chunk_slot(const Pipe& p, int team, int i) {
    return team == 0 ? i + min(i, p.n1) : 2 * i + 1;
}
// Producer: wait until the stage of consumption slot c is free (its previous MMAs have finished reading it).
__device__ __forceinline__ Stage producer_acquire(const Pipe& p, int c) {
    mbar_wait(p.empty0 + 8 * (c % kStages), ((uint32_t)(c / kStages) & 1u) ^ 1u);
    return stage_of(p, c);
}
// Producer: this thread's stores to the stage are done -> visible to the async proxy, count it on the full barrier.
__device__ __forceinline__ void producer_commit(const Pipe& p, int c) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(p.full0 + 8 * (c % kStages));
}

// MMA warp: lane 0 consumes the chunks in slot order.
__device__ __forceinline__ void mma_loop(const Pipe& p, int passes) {
    if ((threadIdx.x & 31) != 0) return;
    const int n = p.n0 + p.n1;
    const uint32_t idesc = idesc_tf32(TM, TN);
    uint32_t acc = 0;
    for (int c = 0; c < n; ++c) {
        const int s = c % kStages;
        mbar_wait(p.full0 + 8 * s, (uint32_t)(c / kStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(p.base + s * kStageBytes), a_lo = a_hi + kTileBytes;
        const uint32_t b_hi = a_hi + 2 * kTileBytes, b_lo = a_hi + 3 * kTileBytes;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
            const uint32_t off = ks * 32;
            const uint64_t da_h = desc_k_sw128(a_hi + off), db_h = desc_k_sw128(b_hi + off);
            if (passes == 3) {
                mma_tf32(p.tmem, desc_k_sw128(a_lo + off), db_h, idesc, acc);
                mma_tf32(p.tmem, da_h, desc_k_sw128(b_lo + off), idesc, 1u);
                acc = 1u;
            }
            mma_tf32(p.tmem, da_h, db_h, idesc, acc);
            acc = 1u;
        }
        umma_commit(p.empty0 + 8 * s);
    }
    if (n > 0) umma_commit(p.done);
}
// Epilogue threads (producer warps): the accumulator is complete and the ring is no longer read.
__device__ __forceinline__ void wait_accumulator(const Pipe& p) {
    if (p.n0 + p.n1 > 0) mbar_wait(p.done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void pipe_teardown(const Pipe& p) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();  // the MMA warp's lane 0 ran alone in mma_loop: converge before the warp-collective dealloc
    if ((threadIdx.x >> 5) == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(p.tmem), "r"((uint32_t)TN) : "memory");
}

// Accumulator row (TMEM lane) 32 * (warp % 4) + lane, columns [c0, c0 + 32). Warp-collective (.sync.aligned): call it
// with the whole warp converged; a warp may only touch its own lane quarter.
__device__ __forceinline__ void load_acc32(const Pipe& p, int c0, float* out) {
    uint32_t v[32];
    const uint32_t taddr = p.tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
}
// Epilogue geometry of producer thread `threadIdx.x` (< 256): accumulator row and the 64-column half it reads.
__device__ __forceinline__ int epi_row() { return ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31); }
__device__ __forceinline__ int epi_col0() { return (threadIdx.x >> 7) * 64; }

// ---------------------------------------------------------------- operand staging (one team: tt = thread in team)
__device__ __forceinline__ void store_split4(uint8_t* hi, uint8_t* lo, uint32_t off, float4 v, int passes) {
    uint4 h;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    *reinterpret_cast<uint4*>(hi + off) = h;
    if (passes == 3) {
        float4 l;
        l.x = v.x - __uint_as_float(h.x); l.y = v.y - __uint_as_float(h.y);
        l.z = v.z - __uint_as_float(h.z); l.w = v.w - __uint_as_float(h.w);
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}
__device__ __forceinline__ void store_split1(uint8_t* hi, uint8_t* lo, uint32_t off, float v, int passes) {
    const uint32_t h = tf32_hi(v);
    *reinterpret_cast<uint32_t*>(hi + off) = h;
    if (passes == 3) *reinterpret_cast<float*>(lo + off) = v - __uint_as_float(h);
}

// tile(r, k) = src[(row0 + r) * ld + k0 + k]  (K contiguous in memory); zero where row0 + r >= n_rows or k0 + k >= K.
// 128-bit loads (8 lanes cover one 128 B row segment), 128-bit conflict-free shared stores.
__device__ __forceinline__ void stage_kcontig(int tt, const float* __restrict__ src, int64_t ld, int row0, int n_rows,
                                              int k0, int K, uint8_t* hi, uint8_t* lo, int passes) {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = i * kTeam + tt;
        const int r = idx >> 3, c4 = idx & 7;
        const int gr = row0 + r, gk = k0 + c4 * 4;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < n_rows) {
            const float* p = src + (int64_t)gr * ld + gk;
            if (vec && gk + 4 <= K) {
                v[i] = ld4(p);
            } else {
                if (gk < K) v[i].x = p[0];
                if (gk + 1 < K) v[i].y = p[1];
                if (gk + 2 < K) v[i].z = p[2];
                if (gk + 3 < K) v[i].w = p[3];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = i * kTeam + tt;
        const int r = idx >> 3, c4 = idx & 7;
        store_split4(hi, lo, (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c4 ^ (r & 7)) << 4)), v[i], passes);
    }
}

// tile(r, k) = src[(k0 + k) * ld + row0 + r]  (the tile's row index is contiguous in memory: a transposing stage).
// A warp loads an 8 (k) x 16 (r) patch per iteration: 64 B runs (two full sectors) per k, and the four scalar stores
// of a lane land in 16 distinct banks across the warp (2-way conflict instead of 16-way for a row-contiguous map).
__device__ __forceinline__ void stage_transposed(int tt, const float* __restrict__ src, int64_t ld, int row0,
                                                 int n_rows, int k0, int K, uint8_t* hi, uint8_t* lo, int passes) {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((row0 & 3) == 0);
    const int warp = tt >> 5, lane = tt & 31;
    const int kq = lane & 7, q = lane >> 3;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int patch = i * 4 + warp;
        const int k = (patch & 3) * 8 + kq, r = (patch >> 2) * 16 + q * 4;
        const int gk = k0 + k, gr = row0 + r;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gk < K) {
            const float* p = src + (int64_t)gk * ld + gr;
            if (vec && gr + 4 <= n_rows) {
                v[i] = ld4(p);
            } else {
                if (gr < n_rows) v[i].x = p[0];
                if (gr + 1 < n_rows) v[i].y = p[1];
                if (gr + 2 < n_rows) v[i].z = p[2];
                if (gr + 3 < n_rows) v[i].w = p[3];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int patch = i * 4 + warp;
        const int k = (patch & 3) * 8 + kq, r = (patch >> 2) * 16 + q * 4;
        // rows r..r+3 share r/8 (r is a multiple of 4) and differ in (r%8): offsets differ by 128 B and by the XOR term
        const uint32_t base = (uint32_t)((r >> 3) * 1024 + ((k & 3) << 2));
        const int kc = k >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r7 = (r + j) & 7;
            const float x = j == 0 ? v[i].x : (j == 1 ? v[i].y : (j == 2 ? v[i].z : v[i].w));
            store_split1(hi, lo, base + r7 * 128 + ((kc ^ r7) << 4), x, passes);
        }
    }
}

// zero one [128 x 32] tile pair (hi and lo): the canvas of a CSR-scattered operand
__device__ __forceinline__ void zero_tiles(int tt, uint8_t* hi, uint8_t* lo, int passes) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < kTileBytes / 16 / kTeam; ++i) {
        reinterpret_cast<uint4*>(hi)[i * kTeam + tt] = z;
        if (passes == 3) reinterpret_cast<uint4*>(lo)[i * kTeam + tt] = z;
    }
}

// first index e in [lo, hi) with a[e] >= key (a ascending)
__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Optional per-row table of CSR windows at 128-column tile boundaries: tab[row * (n_tiles + 1) + j] = first entry of
// the row with column >= 128 * j (j = n_tiles: the row's end). mode 0: no table (binary search), 1: indexed by the
// batch-row index, 2: indexed by the CSR row id.
struct TileTab {
    const int32_t* tab;
    int mode, n_tiles;
};
__device__ __forceinline__ void row_window(const TileTab& tt, const int32_t* __restrict__ indptr,
                                           const int32_t* __restrict__ indices, int batch_row, int u, int tile_lo,
                                           int tile_hi, int& e0, int& s, int& e) {
    e0 = indptr[u];
    if (tt.mode != 0) {
        const int32_t* row = tt.tab + (int64_t)(tt.mode == 1 ? batch_row : u) * (tt.n_tiles + 1);
        s = row[tile_lo];
        e = row[min(tile_hi, tt.n_tiles)];
    } else {
        const int e1 = indptr[u + 1];
        s = lower_bound_i32(indices, e0, e1, tile_lo * TN);
        e = lower_bound_i32(indices, s, e1, tile_hi * TN);
    }
}

}  // namespace umma
}  // namespace dmt
