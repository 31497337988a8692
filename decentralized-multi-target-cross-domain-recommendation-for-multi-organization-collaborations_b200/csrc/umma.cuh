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
// anyway — into the canonical K-major SWIZZLE_128B layout that the UMMA shared-memory descriptor describes
// (CUTLASS cute/atom/mma_traits_sm100.hpp, "Swizzle<3,4,3> o ((8,n),2):((8,SBO),1)" in 16-byte units): a tile is
// 128 rows x 128 B (32 fp32 along K); row r lives at (r/8)*1024 + (r%8)*128, its 16-byte chunk j is stored at chunk
// j ^ (r%8); one tf32 MMA consumes K = 8 elements = 32 B, so the descriptor start address advances by 32 B per k-step
// inside the 128 B swizzle atom.
//
// A kernel is single-buffered on purpose: 64 KB of operand tiles + 128 TMEM columns per CTA let three CTAs share an
// SM, so one CTA's staging overlaps another's MMAs without warp specialisation.
#pragma once
#include "common.cuh"

namespace dmt {
namespace umma {

constexpr int TM = 128, TN = 128, TK = 32;
constexpr int kThreads = 128;
constexpr int kTileBytes = TM * 128;           // one operand, one of hi/lo, one k-chunk: 16 KB
constexpr int kOperandBytes = 4 * kTileBytes;  // A_hi, A_lo, B_hi, B_lo
constexpr int kCtrlBytes = 64;                 // mbarrier + TMEM address slot
// dynamic shared memory of a kernel whose extra (epilogue / index) area needs `extra` bytes beyond the operand tiles
constexpr int smem_bytes(int extra) { return 1024 /*alignment slack*/ + kOperandBytes + extra + kCtrlBytes; }

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
// Bounded: a commit that never arrives (a malformed descriptor, a lost MMA) traps after ~seconds instead of hanging
// the GPU; the host then sees a launch failure.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
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

struct Ctx {
    uint8_t* base;  // 1024 B aligned start of the operand tiles (also the start of a kernel's epilogue area)
    uint8_t *A_hi, *A_lo, *B_hi, *B_lo;
    uint32_t mbar, tmem, parity, accumulate;
};

// Called by all 128 threads, after every early exit. `extra` = the kernel's bytes between operand tiles and control.
__device__ __forceinline__ Ctx setup(uint8_t* raw, int extra) {
    Ctx c;
    c.base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    c.A_hi = c.base;
    c.A_lo = c.base + kTileBytes;
    c.B_hi = c.base + 2 * kTileBytes;
    c.B_lo = c.base + 3 * kTileBytes;
    uint8_t* ctrl = c.base + kOperandBytes + extra;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ctrl);
    uint32_t* slot = reinterpret_cast<uint32_t*>(ctrl + 16);
    c.mbar = smem_u32(mbar);
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
                     "r"((uint32_t)TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tmem = *slot;
    c.parity = 0;
    c.accumulate = 0;
    return c;
}

// All threads, after the chunk's operand tiles were written: make them visible to the tensor core, then one thread
// issues the chunk's MMAs and commits them to the mbarrier.
__device__ __forceinline__ void issue(Ctx& c, int passes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (MMA) reads
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = idesc_tf32(TM, TN);
        uint32_t acc = c.accumulate;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
            const uint32_t off = ks * 32;
            const uint64_t da_h = desc_k_sw128(smem_u32(c.A_hi) + off), db_h = desc_k_sw128(smem_u32(c.B_hi) + off);
            if (passes == 3) {
                const uint64_t da_l = desc_k_sw128(smem_u32(c.A_lo) + off), db_l = desc_k_sw128(smem_u32(c.B_lo) + off);
                mma_tf32(c.tmem, da_l, db_h, idesc, acc);
                mma_tf32(c.tmem, da_h, db_l, idesc, 1u);
                acc = 1u;
            }
            mma_tf32(c.tmem, da_h, db_h, idesc, acc);
            acc = 1u;
        }
        // arrives on the mbarrier once every MMA issued so far has finished reading shared memory / writing TMEM
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.mbar)
                     : "memory");
    }
    c.accumulate = 1u;
}
// All threads: the chunk's MMAs are done — the operand tiles may be overwritten, the accumulator may be read.
__device__ __forceinline__ void wait(Ctx& c) {
    mbar_wait(c.mbar, c.parity);
    c.parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void teardown(Ctx& c) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"((uint32_t)TN) : "memory");
}

// Accumulator row `threadIdx.x` (TMEM lane), columns [c0, c0 + 32): a warp may only touch its own 32 lanes.
__device__ __forceinline__ void load_acc32(const Ctx& c, int c0, float* out) {
    uint32_t v[32];
    const uint32_t taddr = c.tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)c0;
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

// ---------------------------------------------------------------- operand staging (all 128 threads)
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
__device__ __forceinline__ void stage_kcontig(const float* __restrict__ src, int64_t ld, int row0, int n_rows, int k0,
                                              int K, uint8_t* hi, uint8_t* lo, int passes) {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = i * kThreads + threadIdx.x;
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
        const int idx = i * kThreads + threadIdx.x;
        const int r = idx >> 3, c4 = idx & 7;
        store_split4(hi, lo, (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c4 ^ (r & 7)) << 4)), v[i], passes);
    }
}

// tile(r, k) = src[(k0 + k) * ld + row0 + r]  (the tile's row index is contiguous in memory: a transposing stage).
// A warp loads an 8 (k) x 16 (r) patch per iteration: 64 B runs (two full sectors) per k, and the four scalar stores
// of a lane land in 16 distinct banks across the warp (2-way conflict instead of 16-way for a row-contiguous map).
__device__ __forceinline__ void stage_transposed(const float* __restrict__ src, int64_t ld, int row0, int n_rows,
                                                 int k0, int K, uint8_t* hi, uint8_t* lo, int passes) {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((row0 & 3) == 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        store_split1(hi, lo, tile_off(r, k), v[i].x, passes);
        store_split1(hi, lo, tile_off(r + 1, k), v[i].y, passes);
        store_split1(hi, lo, tile_off(r + 2, k), v[i].z, passes);
        store_split1(hi, lo, tile_off(r + 3, k), v[i].w, passes);
    }
}

// zero one [128 x 32] tile pair (hi and lo): the canvas of a CSR-scattered operand
__device__ __forceinline__ void zero_tiles(uint8_t* hi, uint8_t* lo, int passes) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < kTileBytes / 16 / kThreads; ++i) {
        reinterpret_cast<uint4*>(hi)[i * kThreads + threadIdx.x] = z;
        if (passes == 3) reinterpret_cast<uint4*>(lo)[i * kThreads + threadIdx.x] = z;
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

}  // namespace umma
}  // namespace dmt
