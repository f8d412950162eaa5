// Shared device helpers of the fused tile kernel and its binning pass (sm_100a): dp4a tap sums on byte-plane
// coefficients, clip/pack, window starts, mbarrier / TMA wrappers, the resident-tile swizzle.
#pragma once
#include "kernels.cuh"

namespace b200comp {

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One output sample of a pass: sum over NW words of (4 samples) x (4 taps held as three byte planes).
// The planes are chained through the accumulator input -- top plane first, each partial sum moved up one
// byte (PRMT, not a shift-add: IMAD/LEA would compete with dp4a for the FMA-heavy pipe) -- so no separate
// recombination is needed: result = sum(s * k) + 2^21 modulo 2^32, exactly Pillow's int accumulator
// (the rounding term 1 << 21 enters as 32 << 16 in the top plane).
template <int NW>
__device__ __forceinline__ int32_t tap_sum(const uint32_t (&wd)[NW], const uint32_t (&k0)[NW], const uint32_t (&k1)[NW],
                                           const uint32_t (&k2)[NW]) {
    int32_t t = 32;
#pragma unroll
    for (int i = 0; i < NW; ++i) t = dp4a_us(wd[i], k2[i], t);
    uint32_t u = __byte_perm((uint32_t)t, 0u, 0x2104);  // << 8
#pragma unroll
    for (int i = 0; i < NW; ++i) u = dp4a_uu(wd[i], k1[i], u);
    u = __byte_perm(u, 0u, 0x2104);
#pragma unroll
    for (int i = 0; i < NW; ++i) u = dp4a_uu(wd[i], k0[i], u);
    return (int32_t)u;
}

// 4 RGBA pixels -> 4 channel words (byte k of each word = pixel k)
__device__ __forceinline__ void transpose4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t &r, uint32_t &g,
                                           uint32_t &b, uint32_t &a) {
    const uint32_t t01 = __byte_perm(p0, p1, 0x5140);  // p0.b0 p1.b0 p0.b1 p1.b1
    const uint32_t t23 = __byte_perm(p2, p3, 0x5140);
    const uint32_t u01 = __byte_perm(p0, p1, 0x7362);  // p0.b2 p1.b2 p0.b3 p1.b3
    const uint32_t u23 = __byte_perm(p2, p3, 0x7362);
    r = __byte_perm(t01, t23, 0x5410);
    g = __byte_perm(t01, t23, 0x7632);
    b = __byte_perm(u01, u23, 0x5410);
    a = __byte_perm(u01, u23, 0x7632);
}

// Two accumulators -> two clipped bytes in one instruction (I2IP): (hi16 of result) = low 16 bits of `upper`,
// byte 1 = clip8(a1), byte 0 = clip8(a0) (Resample.c clip8: arithmetic shift, clamp to [0, 255]).  Two of them
// pack four samples: pack2(a0, a1, pack2(a2, a3, 0)).
__device__ __forceinline__ uint32_t pack2_clip(int32_t a0, int32_t a1, uint32_t upper) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a1 >> kPrecisionBits), "r"(a0 >> kPrecisionBits), "r"(upper));
    return d;
}

// First source sample of output sample `o` (Resample.c precompute_coeffs: xmin), recomputed with the
// same IEEE double operations as the host table builder (no contraction), so no table lookup is needed.
__device__ __forceinline__ int first_tap(int o, double scale, double support) {
    const double center = __dmul_rn((double)o + 0.5, scale);
    const int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    return max(lo, 0);
}

// ---- TMA / mbarrier helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}

template <int NW>
__device__ __forceinline__ void load_coef_row(const uint32_t *__restrict__ table, int j, uint32_t (&k0)[NW], uint32_t (&k1)[NW],
                                              uint32_t (&k2)[NW]) {
    constexpr int S = coef_row_words(NW);
    const uint4 *row = reinterpret_cast<const uint4 *>(table + (int64_t)j * S);
    uint32_t w[S];
#pragma unroll
    for (int v = 0; v < S / 4; ++v) {
        const uint4 t = __ldg(row + v);
        w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        k0[i] = w[i];
        k1[i] = w[NW + i];
        k2[i] = w[2 * NW + i];
    }
}

// (4*words x 4 channel planes x row quads) box of the prepared cutout -> shared memory; completion on `bar`
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, int x, int y, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
// TMA prefetch of a box into L2 (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const void *tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void *tmap, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(x), "r"(y) : "memory");
}
// plain 1-D bulk copy global -> shared (command blocks); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Word offset of canvas pixel (row r, column x) inside a resident tile buffer.  A tile is two halves of
// 32 pixels x kTileH rows, each exactly what one TMA box with CU_TENSOR_MAP_SWIZZLE_128B leaves in shared
// memory: rows of 128 bytes whose 16-byte chunk index is XORed with (row & 7).  The vertical pass walks rows
// with the lanes of a warp at a fixed column: the swizzle spreads those accesses over 8 banks x 4 words.
__device__ __forceinline__ uint32_t ct_off(int r, int x) {
    return (uint32_t)(((x & 32) ? kTileH * 32 : 0) + (r << 5) + ((((x >> 2) ^ r) & 7) << 2) + (x & 3));
}

}  // namespace b200comp
