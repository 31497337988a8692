// Warp-private row gathers through the bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier).
//
// The gather kernels of the AAE step (decoder SDDMM, the segmented dW4 / dW1 reductions, the encoder SpMM) all walk a
// list of 1 KB fp32 rows picked by an index (reference: the [t x 256] embedding-style gathers of
// src/models/ae.py:102,135-142 and their backward). With plain loads every row in flight costs registers
// (2 x float4 per lane and row), so a warp that wants 8 rows in flight carries 64 registers of landing space and the
// SM holds few such warps. Here the rows land in a shared-memory ring instead: ONE lane issues one 1 KB bulk copy per
// row (SASS UBLKCP), the ring slot's mbarrier counts the bytes, the consumers read the slot with conflict-free 128-bit
// shared loads. Bytes in flight are bounded by shared memory (8 KB per warp), not by registers, so a gather block fits
// 5-6 times per SM and the L2 -> SM path stays busy while the dependent index -> row chains of other warps resolve.
#pragma once
#include "common.cuh"

namespace dmt {
namespace bulk {

constexpr int kRowFloats = 256;             // one gathered row: 256 fp32 = 1 KB (H1 of the AAE, src/utils.py:166-171)
constexpr uint32_t kRowBytes = kRowFloats * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void copy_row(uint32_t dst, const void* src, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(kRowBytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// One warp's ring of SLOTS row slots + SLOTS mbarriers. `phase` (warp-uniform) holds, per slot, the parity its next
// wait uses; it carries over from one gather to the next, so a ring is initialised once per kernel.
template <int SLOTS>
struct WarpRing {
    const float* slots;  // generic address of slot 0
    uint32_t slots_s;    // its shared-window address
    uint32_t bars_s;     // shared-window address of the SLOTS mbarriers (8 B each)
    uint32_t phase;

    __device__ __forceinline__ void issue(int slot, const float* src) {
        const uint32_t bar = bars_s + 8u * (uint32_t)slot;
        mbar_expect_tx(bar, kRowBytes);
        copy_row(slots_s + kRowBytes * (uint32_t)slot, src, bar);
    }
    __device__ __forceinline__ void wait(int slot) {
        mbar_wait(bars_s + 8u * (uint32_t)slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
    }
};

// Call once per kernel by the owning warp's lanes before the block-wide barrier that precedes the first gather.
template <int SLOTS>
__device__ __forceinline__ WarpRing<SLOTS> ring_setup(float* slots, uint64_t* bars, int lane) {
    WarpRing<SLOTS> rg{slots, smem_u32(slots), smem_u32(bars), 0u};
    if (lane < SLOTS) mbar_init(rg.bars_s + 8u * (uint32_t)lane, 1);
    fence_barrier_init();
    return rg;
}

// Stream `cnt` (<= 32) rows base[row_l * 256 ...] through the ring, lane l holding the row index of entry l. Rows are
// handed to `consume(t0, nv, w)` four at a time, in order: w[q][0] / w[q][1] are this lane's float4 slices
// [4*lane, 4*lane+4) and [128 + 4*lane, ...) of entry t0 + q, valid for q < nv (warp-uniform). Up to SLOTS rows are in
// flight; a group's slots are refilled (entries t0 + SLOTS ...) once the warp has consumed the group. The ring is
// drained on return.
template <int SLOTS, class F>
__device__ __forceinline__ void gather_rows(WarpRing<SLOTS>& rg, const float* __restrict__ base, int row_l, int cnt,
                                            int lane, F&& consume) {
    static_assert(SLOTS == 8, "two groups of four rows: one being consumed, one in flight");
    if (lane < SLOTS && lane < cnt) rg.issue(lane, base + (int64_t)row_l * kRowFloats);
    for (int t0 = 0; t0 < cnt; t0 += 4) {
        const int sb = t0 & (SLOTS - 1);
        const int nv = min(4, cnt - t0);
        float4 w[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < nv) {
                rg.wait(sb + q);
                const float* s = rg.slots + (sb + q) * kRowFloats + lane * 4;
                w[q][0] = *reinterpret_cast<const float4*>(s);
                w[q][1] = *reinterpret_cast<const float4*>(s + 128);
            } else {
                w[q][0] = w[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        consume(t0, nv, w);
        const int tn = t0 + SLOTS + (lane & 3);
        const int rn = __shfl_sync(0xffffffffu, row_l, tn & 31);
        __syncwarp();  // every lane has consumed the group: its slots may be overwritten
        if (lane < 4 && tn < cnt) rg.issue(sb + lane, base + (int64_t)rn * kRowFloats);
    }
}

}  // namespace bulk
}  // namespace dmt
