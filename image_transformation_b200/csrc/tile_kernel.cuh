// Fused resample + alpha-over tile kernel (sm_100a), the hot kernel of the path, and the binning pass
// that feeds it.
//
// The kernel is persistent: CTA c walks the 64x32 canvas tiles of command stream c (written by the binning
// kernels below).  A tile lives in shared memory while the placements that touch it are composited in z-order
// (compositor.py:12-21); the binning pass has already dropped the placements that cannot show (transparent
// source patch, or hidden by a later opaque placement that covers the tile).  Per resampled placement:
//   1. patch   one TMA box (words x 4 channel planes x rows) of the PREPARED cutout: premultiplied
//              (Convert.c rgbA2rgba), channel-planar, 4 consecutive pixels of one channel per 32-bit word
//   2. H pass  lane <-> output column; Pillow's 22-bit fixed-point taps are held as three byte planes
//              (k = b0 + 256*b1 + 65536*b2, b2 signed) so four taps cost three dp4a and no byte unpacking;
//              result rounded + clipped to uint8 (ImagingResampleHorizontal_8bpc) and written transposed
//              (4 consecutive ROWS of one channel per word)
//   3. V pass  lane <-> output row, same dp4a scheme (ImagingResampleVertical_8bpc), then un-premultiply
//              (rgba2rgbA) and alpha-over (AlphaComposite.c) onto the resident tile
// The tile is written to HBM once (TMA store).  Arithmetic is integer-exact: dp4a partial sums wrap modulo
// 2^32 and the true accumulator fits in int32, exactly as Pillow's int accumulator.
#pragma once
#include "kernels.cuh"

#ifndef B200COMP_UNCHAINED
#define B200COMP_UNCHAINED 0  // tap_sum: 1 = one accumulator per coefficient plane (more ILP, one more ALU op per sample)
#endif

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
#if B200COMP_UNCHAINED
    // three independent chains (one per plane), recombined with two immediate PRMTs and one three-input add
    int32_t t2 = 32;
    uint32_t u1 = 0u, u0 = 0u;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        t2 = dp4a_us(wd[i], k2[i], t2);
        u1 = dp4a_uu(wd[i], k1[i], u1);
        u0 = dp4a_uu(wd[i], k0[i], u0);
    }
    return (int32_t)(u0 + __byte_perm(u1, 0u, 0x2104) + __byte_perm((uint32_t)t2, 0u, 0x1044));
#endif
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

// Resample.c clip8: arithmetic shift, clamp to [0, 255] (one shift + one min-with-relu)
__device__ __forceinline__ uint32_t clip8i(int32_t v) { return (uint32_t)__vimin_s32_relu(v >> kPrecisionBits, 255); }
// Two accumulators -> two clipped bytes in one instruction (I2IP): (hi16 of result) = low 16 bits of `upper`,
// byte 1 = clip8(a1), byte 0 = clip8(a0).  Two of them pack four samples: pack2(a0, a1, pack2(a2, a3, 0)).
#ifndef B200COMP_CVTPACK
#define B200COMP_CVTPACK 1  // measured 1 % faster than VIMNMX.RELU + PRMT (tools/ab.sh)
#endif
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

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- prepare: RGBA cutout -> premultiplied, channel-planar words --------------------------------
// One launch per plan run converts every distinct cutout the tile kernel resamples into the layout
// its dp4a passes consume: every row becomes four channel planes of w4p words, byte k of word g =
// pixel 4*g+k, colours premultiplied (Convert.c rgbA2rgba); w4p = words per plane, padded to a
// multiple of 4 so plane and row strides are 16-byte multiples (a TMA requirement).  The tile kernel
// then fetches a (words x 4 planes x rows) box per placement with ONE cp.async.bulk.tensor.3d and
// does no per-pixel work before the horizontal pass.  Pixels past the row end are zero.
struct PrepDesc {
    const uint8_t *src;
    uint32_t *dst;    // [sh][4][w4p]; null: only the alpha summary is produced (overlays composited as they are)
    uint32_t *flags;  // [ceil(sh/4)][w4p/4] alpha summary of each 4-row x 16-pixel block (zeroed before the launch):
                      // bit 0 = some alpha != 0, bit 1 = some alpha != 255
    int64_t src_pitch;
    int32_t sw, sh;
    int32_t w4p;      // words per channel plane of a row (multiple of 4)
    int32_t vec_ok;   // src 16-byte aligned with pitch % 16 == 0
};

__global__ void __launch_bounds__(256) prepare_cutouts_kernel(const PrepDesc *__restrict__ descs) {
    const PrepDesc d = descs[blockIdx.y];
    const int64_t total = (int64_t)d.sh * d.w4p;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / d.w4p), g = (int)(i - (int64_t)r * d.w4p);
        const int gx = 4 * g;
        uint32_t *o = d.dst ? d.dst + ((int64_t)r * 4) * d.w4p + g : nullptr;  // null: alpha summary only
        if (gx >= d.sw) {  // padding words
            if (o) { o[0] = 0u; o[d.w4p] = 0u; o[2 * d.w4p] = 0u; o[3 * d.w4p] = 0u; }
            continue;
        }
        const uint8_t *rowp = d.src + (int64_t)r * d.src_pitch + (int64_t)gx * 4;
        uint32_t p0, p1 = 0u, p2 = 0u, p3 = 0u;
        if (d.vec_ok && gx + 3 < d.sw) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rowp));
            p0 = v.x; p1 = v.y; p2 = v.z; p3 = v.w;
        } else {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(rowp);
            p0 = __ldg(q);
            if (gx + 1 < d.sw) p1 = __ldg(q + 1);
            if (gx + 2 < d.sw) p2 = __ldg(q + 2);
            if (gx + 3 < d.sw) p3 = __ldg(q + 3);
        }
        uint32_t R, G, B, A;
        transpose4(p0, p1, p2, p3, R, G, B, A);
        if (o) {
            if (((A ^ (A >> 1)) & 0x7f7f7f7fu) == 0u) {
                // every alpha is 0 or 255: MULDIV255(c, a) is c or 0 -> mask the colours with the alpha bytes
                R &= A; G &= A; B &= A;
            } else {
                transpose4(premultiply_px(p0), premultiply_px(p1), premultiply_px(p2), premultiply_px(p3), R, G, B, A);
            }
            o[0] = R; o[d.w4p] = G; o[2 * d.w4p] = B; o[3 * d.w4p] = A;
        }
        // alpha summary: lets the tile kernel skip fully transparent patches and the alpha plane of
        // fully opaque ones (pixels past the row end count as neither)
        const int nv = min(4, d.sw - gx);
        const uint32_t a_all = nv < 4 ? (A | (0xffffffffu << (8 * nv))) : A;
        const uint32_t bits = (A != 0u ? 1u : 0u) | (a_all != 0xffffffffu ? 2u : 0u);
        if (bits) atomicOr(d.flags + (int64_t)(r >> 2) * (d.w4p >> 2) + (g >> 2), bits);
    }
}

// ---- TMA / mbarrier helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// (words x 4 channel planes x rows) box of the prepared cutout -> shared memory; completion on `bar`
__device__ __forceinline__ void tma_load_patch(void *smem_dst, const void *tmap, int word, int row, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the buffer are done
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(word), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// pull the coefficient rows this lane will need into L1 while the patch is being staged
__device__ __forceinline__ void prefetch_coeffs(const uint32_t *__restrict__ pl, int nw, int n_out, int idx) {
    for (int q = 0; q < 3 * nw; ++q) prefetch_l1(pl + (int64_t)q * n_out + idx);
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Word offset of canvas pixel (row r, column x) inside a resident tile buffer.  A tile is two 32x32-pixel
// halves, each exactly what one TMA box with CU_TENSOR_MAP_SWIZZLE_128B leaves in shared memory: rows of
// 128 bytes whose 16-byte chunk index is XORed with (row & 7).  The vertical pass walks rows with the
// lanes of a warp at a fixed column: the swizzle spreads those 32 accesses over 8 banks x 4 words.
__device__ __forceinline__ uint32_t ct_off(int r, int x) {
    return (uint32_t)(((x & 32) << 5) | (r << 5) | ((((x >> 2) ^ r) & 7) << 2) | (x & 3));
}

// ---- H pass ---------------------------------------------------------------------------------
// I[c][jj][rq]: plane c at I + c*iplane_stride, column pitch IPW words (odd), byte k of word rq =
// intermediate row 4*rq+k (relative to source row 4*rw0).
// P[r][c][wx]: exactly the 3-D TMA box (row pitch PBW words, channel plane pitch PBW/4 words).
template <int NW>
__device__ __forceinline__ void tile_hpass(const uint32_t *__restrict__ P, int PBW,
                                           uint32_t *__restrict__ I, int iplane_stride, int IPW, int NRQ, int cw0,
                                           int ox0, int two, double scale, double support,
                                           const uint32_t *__restrict__ plx, int n_out, int nch) {
    // nch = 3 when every source alpha in the patch is 255: the alpha plane is then 255 after both passes
    // (255 * sum(k) + 2^21 >> 22 == 255 because |sum(k) - 2^22| <= taps) and is not computed
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: row offsets go to uniform registers
    if (warp >= kComputeWarps) return;
    // two column groups of 32: the first (kComputeWarps + 1) / 2 warps take columns 0..31, the rest 32..63
    constexpr int kHalf = (kComputeWarps + 1) / 2;
    const bool two_groups = two > 32;
    const int cg = (two_groups && warp >= kHalf) ? 1 : 0;
    const int wfirst = cg ? warp - kHalf : warp;                                     // this warp's index in its group
    const int rstep = two_groups ? (cg ? kComputeWarps - kHalf : kHalf) : kComputeWarps;  // warps in the group
    const int jj = cg * 32 + lane;
    const unsigned act = __ballot_sync(0xffffffffu, jj < two);  // lanes of this warp that own an output column
    if (jj >= two) return;
    const int j = ox0 + jj;
    const int wbase = (first_tap(j, scale, support) >> 2) - cw0;
    uint32_t k0[NW], k1[NW], k2[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        k0[i] = __ldg(plx + (int64_t)(0 * NW + i) * n_out + j);
        k1[i] = __ldg(plx + (int64_t)(1 * NW + i) * n_out + j);
        k2[i] = __ldg(plx + (int64_t)(2 * NW + i) * n_out + j);
    }
    // unit = row quad: 4 source rows x 32 output columns x all channel planes.  With an alpha plane (nch == 4) that
    // plane goes first: if every alpha the warp read for this unit is 0, the premultiplied colours are 0 as well
    // (rgbA2rgba) and so are their sums -- the three colour planes are stored as zeros without being computed
    // (about a quarter of the units of tiles that straddle a cutout's edge).
    const int plane = PBW >> 2;
    for (int rq = wfirst; rq < NRQ; rq += rstep) {
        const uint32_t *row = P + (rq * 4) * PBW + wbase;
        uint32_t *d = I + jj * IPW + rq;
        if (nch == 4) {
            uint32_t o = 0u, nz = 0u;
            int32_t sv[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                uint32_t wd[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    wd[i] = row[3 * plane + rr * PBW + i];
                    nz |= wd[i];
                }
                sv[rr] = tap_sum<NW>(wd, k0, k1, k2);
#if !B200COMP_CVTPACK
                const uint32_t v = clip8i(sv[rr]);
                if (rr == 0) o = v;
                else if (rr == 1) o = __byte_perm(o, v, 0x3240);
                else if (rr == 2) o = __byte_perm(o, v, 0x3410);
                else o = __byte_perm(o, v, 0x4210);
#endif
            }
#if B200COMP_CVTPACK
            o = pack2_clip(sv[0], sv[1], pack2_clip(sv[2], sv[3], 0u));
#endif
            d[3 * iplane_stride] = o;
            if (!__any_sync(act, nz != 0u)) {
                d[0] = 0u;
                d[iplane_stride] = 0u;
                d[2 * iplane_stride] = 0u;
                continue;
            }
        }
        for (int c = 0; c < 3; ++c) {
            uint32_t o = 0u;
            int32_t sv[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                uint32_t wd[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) wd[i] = row[rr * PBW + i];
                sv[rr] = tap_sum<NW>(wd, k0, k1, k2);
#if !B200COMP_CVTPACK
                const uint32_t v = clip8i(sv[rr]);
                if (rr == 0) o = v;
                else if (rr == 1) o = __byte_perm(o, v, 0x3240);
                else if (rr == 2) o = __byte_perm(o, v, 0x3410);
                else o = __byte_perm(o, v, 0x4210);
#endif
            }
#if B200COMP_CVTPACK
            o = pack2_clip(sv[0], sv[1], pack2_clip(sv[2], sv[3], 0u));
#endif
            *d = o;
            row += plane;
            d += iplane_stride;
        }
    }
}

// ---- V pass + un-premultiply + over -------------------------------------------------------------
// One output column of the tile (lane <-> output row): NCH channel sums, then the pixel goes onto the
// resident tile.  NCH == 3: every source alpha of the column is 255 -- the alpha plane is not computed and
// the pixel replaces the canvas pixel.
template <int NW, int NCH>
__device__ __forceinline__ void vpass_column(const uint32_t *__restrict__ col, int iplane_stride, uint32_t *__restrict__ cpx,
                                             const uint32_t (&k0)[NW], const uint32_t (&k1)[NW], const uint32_t (&k2)[NW]) {
    int32_t acc[4];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t wd[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) wd[i] = col[c * iplane_stride + i];
        acc[c] = tap_sum<NW>(wd, k0, k1, k2);
    }
    // bytes packed with PRMT (the shift-and-or form compiles to IMAD.SHL on the FMA-heavy pipe)
#if B200COMP_CVTPACK
    if (NCH == 3) {
        *cpx = pack2_clip(acc[0], acc[1], pack2_clip(acc[2], 255 << kPrecisionBits, 0u));
    } else {
#else
    const uint32_t rg = __byte_perm(clip8i(acc[0]), clip8i(acc[1]), 0x1140);
    if (NCH == 3) {
        *cpx = __byte_perm(rg, clip8i(acc[2]), 0x5410) | 0xff000000u;
    } else {
#endif
        // Alpha tests on the raw accumulator (clip8(acc) == 0 / == 255).  Do NOT test the clamped value:
        // CUDA 12.9 ptxas folds `clamp(x) == 255` into VIMNMX.RELU's predicate output with the wrong
        // sense on sm_100a (partially transparent pixels took the opaque branch).
        if (acc[3] >= (1 << kPrecisionBits)) {  // else transparent: canvas pixel unchanged
            const bool opaque = acc[3] >= (255 << kPrecisionBits);
#if B200COMP_CVTPACK
            const uint32_t s = pack2_clip(acc[0], acc[1], pack2_clip(acc[2], acc[3], 0u));
#else
            const uint32_t s = __byte_perm(rg, __byte_perm(clip8i(acc[2]), clip8i(acc[3]), 0x1140), 0x5410);
#endif
            *cpx = opaque ? s : over_px(*cpx, unpremultiply_px(s));
        }
    }
}

// NCH == 3: the whole source patch is opaque (alpha summary).  NCH == 4: each column is classified first from
// the alpha plane of the intermediate (one word per lane, two warp reductions): all zero -> nothing to draw,
// the column is skipped; all 255 -> the 3-channel path; on tiles that straddle a cutout's edge about a fifth of
// the columns fall in either class.
template <int NW, int NCH>
__device__ __forceinline__ void tile_vpass_over(const uint32_t *__restrict__ I, int iplane_stride, int IPW, int NRQ,
                                                uint32_t *__restrict__ ctile, int rw0, int oy0, int tho, int two,
                                                int tile_dx, int tile_dy, double scale, double support,
                                                const uint32_t *__restrict__ ply, int n_out) {
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp >= kComputeWarps) return;
    // Lanes past the tile's last row redo the last row (same loads, same value stored to the same address): the
    // warp stays converged (the column classification below is a warp-wide reduction).
    const int lrow = min(lane, tho - 1);
    const int y = oy0 + lrow;
    const int wbase = (first_tap(y, scale, support) >> 2) - rw0;
    uint32_t k0[NW], k1[NW], k2[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        k0[i] = __ldg(ply + (int64_t)(0 * NW + i) * n_out + y);
        k1[i] = __ldg(ply + (int64_t)(1 * NW + i) * n_out + y);
        k2[i] = __ldg(ply + (int64_t)(2 * NW + i) * n_out + y);
    }
    const int r = tile_dy + lrow;
    uint32_t *crow = ctile + (r << 5);
    for (int x = warp; x < two; x += kComputeWarps) {
        const uint32_t *col = I + x * IPW + wbase;
        const int X = tile_dx + x;  // warp-uniform
        uint32_t *cpx = crow + (((X & 32) << 5) | (X & 3)) + ((((X >> 2) ^ r) & 7) << 2);
        if (NCH == 3) {
            vpass_column<NW, 3>(col, iplane_stride, cpx, k0, k1, k2);
        } else {
            const uint32_t *acol = I + 3 * iplane_stride + x * IPW;
            uint32_t any = 0u, all = 0xffffffffu;
            for (int q = lane; q < NRQ; q += 32) {
                const uint32_t a = acol[q];
                any |= a;
                all &= a;
            }
            any = __reduce_or_sync(0xffffffffu, any);
            all = __reduce_and_sync(0xffffffffu, all);
            if (any == 0u) continue;  // every source alpha under this column is 0
            if (all == 0xffffffffu) vpass_column<NW, 3>(col, iplane_stride, cpx, k0, k1, k2);
            else vpass_column<NW, 4>(col, iplane_stride, cpx, k0, k1, k2);
        }
    }
}

struct DevPlacementT {
    const uint8_t *src;    // mode 0: w x h overlay composited as is (raw RGBA)
    const uint32_t *plx;   // [3*nwx][w] coefficient byte planes of the horizontal pass
    const uint32_t *ply;   // [3*nwy][h] vertical pass
    const void *tmap;      // mode 1: CUtensorMap over the prepared cutout, box = (pbw/4 words, 4 planes, nrbox rows)
                           // mode 0: CUtensorMap over the raw overlay, box = 64 x 32 pixels (null: generic loads)
    const uint32_t *flags; // mode 1: alpha summary of the prepared cutout, [sh4][wq] (see PrepDesc)
    double scale_x, support_x;  // sw / w and 3 * max(1, scale): exactly the host builder's doubles
    double scale_y, support_y;
    int32_t src_pitch;     // bytes (mode 0)
    int32_t sw, sh;
    int32_t x, y, w, h;    // destination box
    int32_t nwx, nwy;      // words per output sample (3, 4 or 5)
    int32_t mode;          // 0 = plain over, 1 = resample in the tile kernel
    int32_t pbw;           // TMA box width in words (= 4 * patch words per row)
    int32_t nrbox;         // TMA box height in rows (multiple of 4)
    int32_t wq, sh4;       // alpha summary extent: blocks per row, block rows
};
static_assert(sizeof(DevPlacementT) == 128, "DevPlacementT layout");

// Geometry of a resampled placement on a tile (identical doubles to the host table builder).
struct Geo {
    int ix0, iy0, two, tho, ox0, oy0, cw0, rw0, NRQ, bq0, bq1;
};
__device__ __forceinline__ Geo tile_geometry(const DevPlacementT &d, int tx0, int ty0, int tx1, int ty1) {
    Geo g;
    g.ix0 = max(tx0, d.x);
    g.iy0 = max(ty0, d.y);
    const int ix1 = min(tx1, d.x + d.w), iy1 = min(ty1, d.y + d.h);
    g.two = ix1 - g.ix0;
    g.tho = iy1 - g.iy0;
    g.ox0 = g.ix0 - d.x;
    g.oy0 = g.iy0 - d.y;
    const int w_first = first_tap(g.ox0, d.scale_x, d.support_x) >> 2;
    const int w_last = (first_tap(ix1 - 1 - d.x, d.scale_x, d.support_x) >> 2) + d.nwx - 1;
    g.bq0 = w_first >> 2;  // alpha summary blocks (4 words) the patch touches
    g.bq1 = min(w_last >> 2, d.wq - 1);
    g.cw0 = w_first & ~3;  // TMA boxes start on 16-byte boundaries
    g.rw0 = first_tap(g.oy0, d.scale_y, d.support_y) >> 2;
    g.NRQ = (first_tap(iy1 - 1 - d.y, d.scale_y, d.support_y) >> 2) + d.nwy - g.rw0;
    return g;
}

__device__ __forceinline__ void store_cmd(Cmd *dst, const uint32_t (&w)[16]) {
    uint4 *q = reinterpret_cast<uint4 *>(dst);
    q[0] = make_uint4(w[0], w[1], w[2], w[3]);
    q[1] = make_uint4(w[4], w[5], w[6], w[7]);
    q[2] = make_uint4(w[8], w[9], w[10], w[11]);
    q[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

// ---- binning: (canvases, placements) -> one command stream per persistent CTA -----------------------
// Tile t (numbered over the canvases of the run) belongs to CTA t % G and is its (t / G)-th tile.
//
// count: warp = tile, lane = placement (32 at a time).  Decides which placements become steps of the tile:
// the box must touch it and, for resampled placements, the OR of the alpha summary over the source patch
// must have some non-zero alpha (the whole warp reads the summary rectangle of each candidate).  The keep /
// opaque masks go to `masks` for the fill kernel; the tile's slot count (1 + steps, or 0 for a tile nothing
// is drawn on -- such tiles never enter a stream) is stored stream-major for the scan.
#ifndef B200COMP_BIN_WARPS
#define B200COMP_BIN_WARPS 4
#endif
constexpr int kBinWarps = B200COMP_BIN_WARPS;  // tiles per block (16 measured slower: uneven tiles hold block slots)
#ifndef B200COMP_BIN_MINBLOCKS
#define B200COMP_BIN_MINBLOCKS 10  // resident blocks per SM the binning kernels are compiled for (48 registers: 11 % faster than uncapped)
#endif
__global__ void __launch_bounds__(kBinWarps * 32, B200COMP_BIN_MINBLOCKS)
bin_count_kernel(const DevCanvas *__restrict__ canvases, const DevPlacementT *__restrict__ placements,
                 const int4 *__restrict__ boxes, int64_t run_tile_base, int G, int K, int32_t *__restrict__ cnt, uint32_t *__restrict__ masks,
                 int mask_chunks, int patch_words, int inter_words, unsigned long long *__restrict__ cursor,
                 int *__restrict__ status, int cull) {
    if (cursor && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *cursor = 0ull;  // first launch of a run
    const DevCanvas &cv = canvases[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int local = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
    if (local >= cv.tiles_x * cv.tiles_y) return;
    const int ty = local / cv.tiles_x, tx = local - ty * cv.tiles_x;
    const int tx0 = tx * kTileW, ty0 = ty * kTileH;
    const int tx1 = min(cv.W, tx0 + kTileW), ty1 = min(cv.H, ty0 + kTileH);
    const int64_t t = cv.tile_base - run_tile_base + local;
    const unsigned tu = (unsigned)t;  // tiles of a run fit 31 bits (checked by the host): 32-bit division below
    uint32_t *mk = masks + t * (int64_t)mask_chunks * 2;
    int n = 0;
    int occ_chunk = -1, occ_top = 0;  // last placement that hides everything under it on this tile
    for (int chunk = 0; chunk < mask_chunks; ++chunk) {
        const int i = chunk * 32 + lane;
        if (chunk * 32 >= cv.count) {  // canvases with fewer placements than the widest one of the plan
            if (lane == 0) mk[2 * chunk] = mk[2 * chunk + 1] = 0u;
            continue;
        }
        const DevPlacementT &d = placements[cv.first + min(i, cv.count - 1)];
        const int4 bx = __ldg(boxes + cv.first + min(i, cv.count - 1));  // (x, y, w, h): one 16-byte load per lane
        const bool hit = i < cv.count && max(tx0, bx.x) < min(tx1, bx.x + bx.z) && max(ty0, bx.y) < min(ty1, bx.y + bx.w);
        const int mode = hit ? d.mode : 0;
        // rectangle of the alpha summary (4-row x 16-pixel blocks of the SOURCE) under this tile, and the part of the
        // tile the placement covers
        bool scan = false;
        const uint32_t *s_fl = nullptr;
        int s_wq = 0, s_c0 = 0, s_nbw = 0, s_r0 = 0, s_r1 = 0, two = 0, tho = 0;
        if (hit && mode != 0) {
            const Geo g = tile_geometry(d, tx0, ty0, tx1, ty1);
            const bool fits = d.pbw * d.nrbox <= patch_words && 4 * kTileW * (g.NRQ | 1) <= inter_words &&
                              4 * g.NRQ <= d.nrbox && g.cw0 < 65536 && g.rw0 < 65536;
            if (!fits) atomicOr(status, kStatusPatchOverflow);  // host sizing bug: flagged, step dropped
            scan = fits;
            s_fl = d.flags;
            s_wq = d.wq;
            s_c0 = g.bq0;
            s_nbw = g.bq1 - g.bq0 + 1;
            s_r0 = g.rw0;
            s_r1 = min(g.rw0 + g.NRQ, d.sh4);
            two = g.two;
            tho = g.tho;
        } else if (hit) {  // overlay composited as it is: source pixel = canvas pixel - box origin
            const int ix0 = max(tx0, bx.x), iy0 = max(ty0, bx.y), ix1 = min(tx1, bx.x + bx.z), iy1 = min(ty1, bx.y + bx.w);
            two = ix1 - ix0;
            tho = iy1 - iy0;
            s_fl = d.flags;
            scan = s_fl != nullptr;
            s_wq = d.wq;
            s_c0 = (ix0 - bx.x) >> 4;
            s_nbw = ((ix1 - 1 - bx.x) >> 4) - s_c0 + 1;
            s_r0 = (iy0 - bx.y) >> 2;
            s_r1 = ((iy1 - 1 - bx.y) >> 2) + 1;
        }
        uint32_t my_bits = 0u;
        for (uint32_t mr = __ballot_sync(0xffffffffu, scan); mr; mr &= mr - 1u) {
            const int b = __ffs((int)mr) - 1;
            const unsigned long long fp = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(s_fl), b);
            const int wq = __shfl_sync(0xffffffffu, s_wq, b), bq0 = __shfl_sync(0xffffffffu, s_c0, b);
            const int nbw = __shfl_sync(0xffffffffu, s_nbw, b);
            const int r0 = __shfl_sync(0xffffffffu, s_r0, b), r1 = __shfl_sync(0xffffffffu, s_r1, b);
            const uint32_t *fl = reinterpret_cast<const uint32_t *>((uintptr_t)fp);
            const int nb = nbw * (r1 - r0);
            uint32_t bits = 0u;
            for (int q = lane; q < nb; q += 32) {
                const int br = q / nbw, bc = q - br * nbw;
                bits |= __ldg(fl + (int64_t)(r0 + br) * wq + bq0 + bc);
            }
            bits = __reduce_or_sync(0xffffffffu, bits);
            if (lane == b) my_bits = bits;
        }
        // keep: unless nothing but alpha 0 lies under the tile (overlays without a summary are always kept;
        // resampled placements that do not fit the kernel's buffers were flagged above and are dropped)
        const bool keep = hit && (scan ? (my_bits & 1u) != 0u : mode == 0);
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        const bool opaque = keep && scan && !(my_bits & 2u);  // every alpha 255
        const uint32_t om = __ballot_sync(0xffffffffu, opaque);
        // Occlusion: an opaque placement whose box covers the whole tile replaces every pixel of it (the V pass
        // stores its pixels without reading the canvas; over_px returns an opaque source pixel as it is), so
        // nothing drawn earlier -- earlier placements and the background -- can show.  Those steps are dropped; the result is unchanged bit for bit.
        const uint32_t cm = __ballot_sync(0xffffffffu, cull && opaque && two == tx1 - tx0 && tho == ty1 - ty0);
        if (lane == 0) {
            mk[2 * chunk] = km;
            mk[2 * chunk + 1] = om;
        }
        if (cm) {
            occ_chunk = chunk;
            occ_top = 31 - __clz((int)cm);
            n = __popc(km >> occ_top);
        } else {
            n += __popc(km);
        }
    }
    if (lane == 0) {
        if (occ_chunk >= 0) {
            for (int c = 0; c < occ_chunk; ++c) mk[2 * c] = mk[2 * c + 1] = 0u;
            mk[2 * occ_chunk] &= ~((1u << occ_top) - 1u);
            mk[2 * occ_chunk + 1] &= ~((1u << occ_top) - 1u);
        }
        cnt[(int64_t)(tu % (unsigned)G) * K + tu / (unsigned)G] = n ? n + 1 : 0;
    }
}

// exclusive scan of every stream's row (in place), one warp per stream; the stream's region of the record
// array is claimed with one atomicAdd (streams need not be stored in order), END record written.
__global__ void __launch_bounds__(256)
bin_scan_kernel(int32_t *__restrict__ cnt, int G, int K, int64_t n_tiles, int64_t *__restrict__ stream_off,
                unsigned long long *__restrict__ cursor, Cmd *__restrict__ streams, int64_t capacity,
                int *__restrict__ status) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= G) return;
    const int64_t nk = n_tiles > c ? (n_tiles - c + G - 1) / G : 0;
    int32_t *row = cnt + (int64_t)c * K;
    int32_t carry = 0;
    for (int64_t k0 = 0; k0 < nk; k0 += 32) {
        const int64_t k = k0 + lane;
        const int32_t v = k < nk ? row[k] : 0;
        int32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (k < nk) row[k] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) {
        const int64_t len = (int64_t)carry + 1;  // + END
        const int64_t base = (int64_t)atomicAdd(cursor, (unsigned long long)len);
        stream_off[c] = base;
        if (base + len <= capacity) streams[base + len - 1].w[0] = kCmdEnd;
        else atomicOr(status, kStatusStreamOverflow);
    }
}

// fill: warp = tile, lane = placement.  Tiles without steps are finished right here (background or solid
// colour copied to the output: they never reach the tile kernel); the others get their TILE record and one
// record per kept placement, in z-order (ballot ranks).
__global__ void __launch_bounds__(kBinWarps * 32, B200COMP_BIN_MINBLOCKS)
bin_fill_kernel(const DevCanvas *__restrict__ canvases, int canvas0, const DevPlacementT *__restrict__ placements,
                int64_t run_tile_base, int G, int K, const int32_t *__restrict__ scan, const uint32_t *__restrict__ masks,
                int mask_chunks, const int64_t *__restrict__ stream_off, Cmd *__restrict__ streams, int64_t capacity,
                const uint8_t *__restrict__ maps_base, const uint32_t *__restrict__ tables_base) {
    const DevCanvas &cv = canvases[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int local = blockIdx.x * kBinWarps + (threadIdx.x >> 5);
    if (local >= cv.tiles_x * cv.tiles_y) return;
    const int ty = local / cv.tiles_x, tx = local - ty * cv.tiles_x;
    const int tx0 = tx * kTileW, ty0 = ty * kTileH;
    const int tx1 = min(cv.W, tx0 + kTileW), ty1 = min(cv.H, ty0 + kTileH);
    const int64_t t = cv.tile_base - run_tile_base + local;
    const uint32_t *mk = masks + t * (int64_t)mask_chunks * 2;
    int n_steps = 0;
    for (int c = 0; c < mask_chunks; ++c) n_steps += __popc(mk[2 * c]);
    if (n_steps == 0) {
        // nothing is drawn on this tile: out = background (or the solid colour)
        const int tw = tx1 - tx0, th = ty1 - ty0;
        const bool vec = tw == kTileW && ((reinterpret_cast<uintptr_t>(cv.out) | (uintptr_t)cv.out_pitch) & 15u) == 0 &&
                         (!cv.bg || ((reinterpret_cast<uintptr_t>(cv.bg) | (uintptr_t)cv.bg_pitch) & 15u) == 0);
        if (vec) {  // 16 lanes x 16 bytes per row, two rows per sweep, eight sweeps of loads in flight before the stores
            const int x4 = (lane & 15) * 4;
            const uint4 sv = make_uint4(cv.solid, cv.solid, cv.solid, cv.solid);
            const uint8_t *bg = cv.bg;
            uint8_t *out = cv.out;
            const int64_t bgp = cv.bg_pitch, outp = cv.out_pitch;
            const int64_t px = (int64_t)(tx0 + x4) * 4;
            for (int y0 = lane >> 4; y0 < th; y0 += 16) {
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int yy = y0 + 2 * k;
                    v[k] = (bg && yy < th) ? __ldg(reinterpret_cast<const uint4 *>(bg + (int64_t)(ty0 + yy) * bgp + px)) : sv;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int yy = y0 + 2 * k;
                    if (yy < th) *reinterpret_cast<uint4 *>(out + (int64_t)(ty0 + yy) * outp + px) = v[k];
                }
            }
        } else {
            for (int yy = 0; yy < th; ++yy)
                for (int xx = lane; xx < tw; xx += 32) {
                    const int64_t px = (int64_t)(tx0 + xx) * 4;
                    const uint32_t v = cv.bg ? ld_px(cv.bg, (int64_t)(ty0 + yy) * cv.bg_pitch + px) : cv.solid;
                    *reinterpret_cast<uint32_t *>(cv.out + (int64_t)(ty0 + yy) * cv.out_pitch + px) = v;
                }
        }
        return;
    }
    const unsigned tu = (unsigned)t;  // tiles of a run fit 31 bits (checked by the host): 32-bit division
    const int64_t base = stream_off[tu % (unsigned)G] + scan[(int64_t)(tu % (unsigned)G) * K + tu / (unsigned)G];
    int n_slots = 0;
    uint32_t w[16];
    bool first_seen = false;
    uint32_t nobg = 0u;  // the tile's first step replaces every pixel: the background is never read
    for (int i0 = 0, chunk = 0; i0 < cv.count && chunk < mask_chunks; i0 += 32, ++chunk) {
        const uint32_t km = mk[2 * chunk], om = mk[2 * chunk + 1];
        bool occludes = false;
        if ((km >> lane) & 1u) {
            const DevPlacementT &d = placements[cv.first + i0 + lane];
            const int64_t at = base + 1 + n_slots + __popc(km & ((1u << lane) - 1u));
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = 0u;
            if (d.mode == 0) {
                const int ix0 = max(tx0, d.x), iy0 = max(ty0, d.y);
                const int two = min(tx1, d.x + d.w) - ix0, tho = min(ty1, d.y + d.h) - iy0;
                w[2] = (uint32_t)(ix0 - tx0) | ((uint32_t)(iy0 - ty0) << 8) | ((uint32_t)two << 16) | ((uint32_t)tho << 24);
                occludes = ((om >> lane) & 1u) && two == tx1 - tx0 && tho == ty1 - ty0;
                if (d.tmap) {
                    // TMA boxes start on 16-byte boundaries: the box begins up to 3 pixels left of the tile
                    // origin (w5 = that shift) and is kOverlayBoxW = 68 pixels wide
                    w[0] = kCmdIdentTma;
                    const int cx = tx0 - d.x;
                    w[3] = (uint32_t)(cx & ~3);
                    w[5] = (uint32_t)(cx & 3);
                    w[4] = (uint32_t)(ty0 - d.y);
                    w[8] = (uint32_t)((reinterpret_cast<const uint8_t *>(d.tmap) - maps_base) >> 7);
                } else {
                    w[0] = kCmdIdentLdg;
                    w[3] = (uint32_t)(ix0 - d.x);
                    w[4] = (uint32_t)(iy0 - d.y);
                    w[6] = (uint32_t)d.src_pitch;
                    const uint64_t p = reinterpret_cast<uint64_t>(d.src);
                    w[12] = (uint32_t)p;
                    w[13] = (uint32_t)(p >> 32);
                }
            } else {
                const Geo g = tile_geometry(d, tx0, ty0, tx1, ty1);
                occludes = ((om >> lane) & 1u) && g.two == tx1 - tx0 && g.tho == ty1 - ty0;
                w[0] = kCmdResample;
                w[1] = (uint32_t)d.nwx | ((uint32_t)d.nwy << 8) | (((om >> lane) & 1u) ? (3u << 16) : (4u << 16)) | ((uint32_t)g.NRQ << 24);
                w[2] = (uint32_t)(g.ix0 - tx0) | ((uint32_t)(g.iy0 - ty0) << 8) | ((uint32_t)g.two << 16) | ((uint32_t)g.tho << 24);
                w[3] = (uint32_t)g.ox0;
                w[4] = (uint32_t)g.oy0;
                w[5] = (uint32_t)g.cw0 | ((uint32_t)g.rw0 << 16);
                w[6] = (uint32_t)d.w;
                w[7] = (uint32_t)d.h;
                w[8] = (uint32_t)((reinterpret_cast<const uint8_t *>(d.tmap) - maps_base) >> 7);
                w[9] = (uint32_t)(d.plx - tables_base);
                w[10] = (uint32_t)(d.ply - tables_base);
                w[11] = (uint32_t)d.pbw | ((uint32_t)d.nrbox << 16);
                const uint64_t sx = (uint64_t)__double_as_longlong(d.scale_x), sy = (uint64_t)__double_as_longlong(d.scale_y);
                w[12] = (uint32_t)sx; w[13] = (uint32_t)(sx >> 32);
                w[14] = (uint32_t)sy; w[15] = (uint32_t)(sy >> 32);
            }
            if (at < capacity) store_cmd(streams + at, w);  // overflow is flagged by the scan kernel
        }
        if (!first_seen && km) {
            first_seen = true;
            nobg = (__ballot_sync(0xffffffffu, occludes) >> (__ffs((int)km) - 1)) & 1u;
        }
        n_slots += __popc(km);
    }
    if (lane != 0 || base >= capacity) return;
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0u;
    w[0] = kCmdTile;
    w[1] = (uint32_t)n_steps;
    w[2] = (uint32_t)tx0;
    w[3] = (uint32_t)ty0;
    w[4] = (uint32_t)(tx1 - tx0) | ((uint32_t)(ty1 - ty0) << 16);
    w[5] = cv.solid;
    w[6] = nobg ? (kTileNoBg | (cv.out_map ? kTileOutTma : 0u))
                : ((cv.bg ? kTileHasBg : 0u) | (cv.bg && cv.bg_map ? kTileBgTma : 0u) | (cv.out_map ? kTileOutTma : 0u));
    w[7] = (uint32_t)(canvas0 + (int)blockIdx.y);
    const uint64_t bm = reinterpret_cast<uint64_t>(cv.bg_map), om = reinterpret_cast<uint64_t>(cv.out_map);
    w[8] = (uint32_t)bm; w[9] = (uint32_t)(bm >> 32);
    w[10] = (uint32_t)om; w[11] = (uint32_t)(om >> 32);
    store_cmd(streams + base, w);
}

// Optional phase timing (-DB200COMP_PROFILE=1): cycles one observer thread per CTA spends in each phase of
// the step loop, summed over CTAs into g_prof (read with b200comp_debug_profile_).
#ifndef B200COMP_PROFILE
#define B200COMP_PROFILE 0
#endif
#if B200COMP_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_MARK(slot)                                                      \
    do {                                                                     \
        if (tid == B200COMP_PROFILE_TID) {                                   \
            const long long now__ = clock64();                               \
            prof_acc[slot] += (unsigned long long)(now__ - prof_t);          \
            prof_t = now__;                                                  \
        }                                                                    \
    } while (0)
#ifndef B200COMP_PROFILE_TID
#define B200COMP_PROFILE_TID 32
#endif
#else
#define PROF_MARK(slot) do { } while (0)
#endif

// ---- the persistent tile kernel ------------------------------------------------------------------
// CTA c consumes command stream c.  Everything it touches arrives asynchronously and ahead of use:
//   * command records: cp.async into an 8-slot ring, 6 records ahead
//   * background tiles: TMA (two 32x32-pixel boxes, 128-byte swizzle) into one of kTileBufs resident
//     tile buffers, up to two tiles ahead; finished tiles leave through TMA stores (bulk groups)
//   * source patches: one TMA box per step, issued as soon as the previous step's horizontal pass has
//     released the patch buffer -- also across tile boundaries
// Lane 0 of the last warp is the producer (it only issues copies; it never waits for data on behalf of others);
// the other eleven warps compute.
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
composite_stream_kernel(const Cmd *__restrict__ streams, const int64_t *__restrict__ stream_off,
                        const DevCanvas *__restrict__ canvases, const uint8_t *__restrict__ maps,
                        const uint32_t *__restrict__ tables, int patch_words, int inter_words) {
    extern __shared__ uint32_t smem_raw[];
    // tile buffers need 1024-byte alignment (swizzle atom); the launch reserves the slack
    uint32_t *ctile = reinterpret_cast<uint32_t *>(
        reinterpret_cast<uint8_t *>(smem_raw) + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    uint32_t *P = ctile + kTileBufs * kTileWords;  // patch_words (multiple of 32 words: stays 128-byte aligned)
    uint32_t *I = P + patch_words;                 // inter_words
    Cmd *ring = reinterpret_cast<Cmd *>(I + inter_words);
    uint64_t *bg_full = reinterpret_cast<uint64_t *>(ring + kRing);
    uint64_t *patch_full = bg_full + kTileBufs;
    // uniform state kept in shared memory to save registers: the current TILE record, the producer's cursor
    uint32_t *trec2 = reinterpret_cast<uint32_t *>(patch_full + 1);  // 2 x 16 words, by tile parity (the finished
                                                                     // tile is flushed while the next one begins)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ptid = tid - kProducerTid;  // >= 0: producer warp
    const Cmd *stream = streams + stream_off[blockIdx.x];

    if (tid == kProducerTid) {
        for (int b = 0; b < kTileBufs; ++b) mbar_init(&bg_full[b], 1);
        mbar_init(patch_full, 1);
    }
    int fetched = 0;  // records whose copy into the ring has been issued (one cp.async group each)

    // consumer state (uniform across the CTA)
    int pos = 0;          // record being consumed
    int ctseq = -1;       // sequence number of the current tile (buffer ctseq % kTileBufs)
    uint32_t pseq = 0;    // patches consumed so far (parity of patch_full)
    int steps_left = 0;
    bool bg_pending = false;
    uint32_t c_flags = 0;
#define trec (trec2 + ((ctseq & 1) << 4))
#define c_tx0 ((int)trec[2])
#define c_ty0 ((int)trec[3])
#define c_tw ((int)(trec[4] & 0xffffu))
#define c_th ((int)(trec[4] >> 16))
#define c_canvas (trec[7])
#define c_out_map (reinterpret_cast<const void *>((uint64_t)trec[10] | ((uint64_t)trec[11] << 32)))
    // producer state (meaningful in the producer thread only)
    int ppos = 0;             // next record to examine
    int ptseq = 0;            // tiles whose background has been issued
    bool patch_busy = false;  // the patch buffer holds (or is receiving) a patch whose H pass has not finished
    bool p_done = false;      // END seen

    auto producer_advance = [&](int limit) {
        while (!p_done && ppos <= limit) {
            const Cmd &c = ring[ppos & (kRing - 1)];
            const uint32_t kind = c.w[0];
            if (kind == kCmdTile) {
                if (ptseq - ctseq > kTileBufs - 2) break;  // no free tile buffer yet
                if (c.w[6] & kTileBgTma) {
                    const int b = ptseq & (kTileBufs - 1);
                    const int tw = (int)(c.w[4] & 0xffffu);
                    const void *map = reinterpret_cast<const void *>((uint64_t)c.w[8] | ((uint64_t)c.w[9] << 32));
                    bulk_wait_read<1>();  // the store that last read this buffer (kTileBufs tiles ago) is done
                    fence_async_smem();
                    mbar_expect_tx(&bg_full[b], tw > 32 ? 8192u : 4096u);
                    tma_load_2d(ctile + b * kTileWords, map, (int)c.w[2], (int)c.w[3], &bg_full[b]);
                    if (tw > 32) tma_load_2d(ctile + b * kTileWords + 1024, map, (int)c.w[2] + 32, (int)c.w[3], &bg_full[b]);
                } else if (c.w[6] & kTileNoBg) {
                    bulk_wait_read<1>();  // nothing to load, but the buffer must be free before the first step writes it
                }
                ++ptseq;
            } else if (kind == kCmdResample) {
                if (patch_busy) break;
                mbar_expect_tx(patch_full, (c.w[11] & 0xffffu) * (c.w[11] >> 16) * 4u);
                tma_load_patch(P, maps + ((uint64_t)c.w[8] << 7), (int)(c.w[5] & 0xffffu), 4 * (int)(c.w[5] >> 16), patch_full);
                patch_busy = true;
            } else if (kind == kCmdIdentTma) {
                if (patch_busy) break;
                fence_async_smem();
                mbar_expect_tx(patch_full, (uint32_t)(kOverlayBoxW * kTileH) * 4u);
                tma_load_2d(P, maps + ((uint64_t)c.w[8] << 7), (int)c.w[3], (int)c.w[4], patch_full);
                patch_busy = true;
            } else if (kind == kCmdEnd) {
                p_done = true;
                break;
            }
            ++ppos;
        }
    };

    // The tile's pixels are final.  The store itself is issued after the next (A) barrier (flush_tile), so a
    // finished tile costs no barrier of its own.
    // Background loads complete phases of bg_full[b] one by one, but only tiles that HAVE a TMA background use
    // one (solid-colour, generically loaded and fully occluded tiles do not): the parity to wait for is counted
    // per buffer, not derived from the tile sequence number.
    uint32_t bg_phase = 0u;  // bit b = parity of the next phase of bg_full[b]
    auto wait_bg = [&]() {
        const int b = ctseq & (kTileBufs - 1);
        mbar_wait(&bg_full[b], (bg_phase >> b) & 1u);
        bg_phase ^= 1u << b;
        bg_pending = false;
    };
    bool store_pending = false;
    auto finish_tile = [&]() {
        if (bg_pending) {
            wait_bg();
        }
        if (c_flags & kTileOutTma) fence_async_smem();  // generic writes to the tile -> visible to the async proxy
        store_pending = true;
    };
    // after a barrier: write the finished tile out (TMA store by the producer thread, or generic stores by everyone)
    auto flush_tile = [&]() {
        store_pending = false;
        const uint32_t *ct = ctile + (ctseq & (kTileBufs - 1)) * kTileWords;
        if (c_flags & kTileOutTma) {
            if (tid == kProducerTid) {
                tma_store_2d(c_out_map, c_tx0, c_ty0, ct);
                if (c_tw > 32) tma_store_2d(c_out_map, c_tx0 + 32, c_ty0, ct + 1024);
                bulk_commit();
            }
        } else {
            const DevCanvas &cv = canvases[c_canvas];
            uint8_t *out = cv.out;
            const int64_t pitch = cv.out_pitch;
            const int xx = tid & (kTileW - 1);
            if (tid < kElemThreads && xx < c_tw)
                for (int yy = tid / kTileW; yy < c_th; yy += kRowSweep)
                    *reinterpret_cast<uint32_t *>(out + (int64_t)(c_ty0 + yy) * pitch + (int64_t)(c_tx0 + xx) * 4) = ct[ct_off(yy, xx)];
            if (tid == kProducerTid) bulk_commit();  // every tile is one bulk group, so the group counting stays uniform
        }
    };

#if B200COMP_PROFILE
    unsigned long long prof_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = clock64();
#endif
    for (;;) {
        PROF_MARK(0);  // end of the previous record (tile begin / step tail / NOP)
        if (ptid >= 0 && ptid < 4) {
            // top the ring up to record pos + kRingAhead - 1 (an iteration consumes one or two records), then
            // wait until all but the newest kRingAhead - 1 - kLook copies have landed: records <= pos + kLook
            for (; fetched < pos + kRingAhead; ++fetched) {
                cp_async16(reinterpret_cast<uint8_t *>(ring + (fetched & (kRing - 1))) + 16 * ptid,
                           reinterpret_cast<const uint8_t *>(stream + fetched) + 16 * ptid);
                cp_async_commit();
            }
            cp_async_wait<kRingAhead - 1 - kLook>();
        }
        __syncthreads();  // (A) ring visible; every thread is done with the previous record
        PROF_MARK(1);  // barrier (A)
        if (store_pending) flush_tile();
        if (ring[pos & (kRing - 1)].w[0] == kCmdEnd) break;
        const int limit = pos + kLook;  // last record the ring is guaranteed to hold during this iteration
        if (tid == kProducerTid && (ring[pos & (kRing - 1)].w[0] == kCmdTile || ppos <= pos)) producer_advance(limit);
        if (ring[pos & (kRing - 1)].w[0] == kCmdTile) {
            // a TILE record is always followed by its first step: both are consumed in this iteration
            const Cmd &tc = ring[pos & (kRing - 1)];
            ++ctseq;
            steps_left = (int)tc.w[1];
            c_flags = tc.w[6];
            if (tid < 16) trec[tid] = tc.w[tid];  // read after the next barrier at the earliest
            if (c_flags & kTileBgTma) {
                bg_pending = true;
            } else if (c_flags & kTileNoBg) {
                // the first step is an opaque placement over the whole tile: it stores every pixel (after barrier
                // (B), i.e. after the producer has seen this record and waited for the buffer's last store)
                bg_pending = false;
            } else {
                // solid colour, or a background TMA cannot address: fill the buffer here
                const uint32_t solid = tc.w[5];
                if (tid == kProducerTid) bulk_wait_read<kTileBufs - 1>();  // the store that last read this buffer is done
                __syncthreads();
                uint32_t *ctb = ctile + (ctseq & (kTileBufs - 1)) * kTileWords;
                const int xx = tid & (kTileW - 1);
                if (c_flags & kTileHasBg) {
                    const DevCanvas &cv = canvases[c_canvas];
                    const uint8_t *bg = cv.bg;
                    const int64_t pitch = cv.bg_pitch;
                    if (tid < kElemThreads && xx < c_tw)
                        for (int yy = tid / kTileW; yy < c_th; yy += kRowSweep)
                            ctb[ct_off(yy, xx)] = ld_px(bg, (int64_t)(c_ty0 + yy) * pitch + (int64_t)(c_tx0 + xx) * 4);
                } else {
                    if (tid < kElemThreads)
                        for (int yy = tid / kTileW; yy < kTileH; yy += kRowSweep) ctb[ct_off(yy, xx)] = solid;
                }
                __syncthreads();  // the tile's first step (same iteration) may composite right away
                bg_pending = false;
            }
            PROF_MARK(7);  // tile begin
            ++pos;
        }
        const Cmd &cmd = ring[pos & (kRing - 1)];
        const uint32_t kind = cmd.w[0];
        uint32_t *ct = ctile + (ctseq & (kTileBufs - 1)) * kTileWords;
        // Record fields are the same for every thread, but the compiler cannot know that of a shared-memory load:
        // broadcasting them from lane 0 marks them warp-uniform, so the passes' row / column / plane offsets are
        // computed once per warp on the uniform datapath (LDS [R + UR + imm]) instead of per lane with IMADs that
        // compete with dp4a for the FMA-heavy pipe.
        const uint32_t w1 = uni(cmd.w[1]), w2 = uni(cmd.w[2]);
        const int dx = (int)(w2 & 0xffu), dy = (int)((w2 >> 8) & 0xffu);
        const int two = (int)((w2 >> 16) & 0xffu), tho = (int)(w2 >> 24);
        if (kind == kCmdResample) {
            const int nwx = (int)(w1 & 0xffu), nwy = (int)((w1 >> 8) & 0xffu);
            const int nch = (int)((w1 >> 16) & 0xffu), NRQ = (int)(w1 >> 24);
            const uint32_t w5 = uni(cmd.w[5]);
            const int ox0 = (int)uni(cmd.w[3]), oy0 = (int)uni(cmd.w[4]);
            const int cw0 = (int)(w5 & 0xffffu), rw0 = (int)(w5 >> 16);
            const int n_out_x = (int)uni(cmd.w[6]), n_out_y = (int)uni(cmd.w[7]);
            const uint32_t *plx = tables + uni(cmd.w[9]), *ply = tables + uni(cmd.w[10]);
            const int pbw = (int)(uni(cmd.w[11]) & 0xffffu);
            const double scale_x = __longlong_as_double((long long)((uint64_t)cmd.w[12] | ((uint64_t)cmd.w[13] << 32)));
            const double scale_y = __longlong_as_double((long long)((uint64_t)cmd.w[14] | ((uint64_t)cmd.w[15] << 32)));
            // a skipped pass is the 1-tap identity: scale 1, support 1 -> first tap = the sample itself
            const double support_x = scale_x == 1.0 ? 1.0 : __dmul_rn(3.0, fmax(scale_x, 1.0));
            const double support_y = scale_y == 1.0 ? 1.0 : __dmul_rn(3.0, fmax(scale_y, 1.0));
            const int IPW = NRQ | 1;
            const int iplane_stride = kTileW * IPW;
            // coefficient rows -> L1 while the patch is in flight
            if (warp == kComputeWarps) {  // the producer warp's idle lanes do the prefetching
                for (int jj = lane; jj < two; jj += 32) prefetch_coeffs(plx, nwx, n_out_x, ox0 + jj);
                if (lane < tho) prefetch_coeffs(ply, nwy, n_out_y, oy0 + lane);
            }
            PROF_MARK(2);  // dispatch, decode, producer thread, coefficient prefetch
            mbar_wait(patch_full, pseq & 1u);  // source patch has landed in P
            PROF_MARK(3);  // wait for the patch
            ++pseq;
            if (nwx == 3)
                tile_hpass<3>(P, pbw, I, iplane_stride, IPW, NRQ, cw0, ox0, two, scale_x, support_x, plx, n_out_x, nch);
            else if (nwx == 4)
                tile_hpass<4>(P, pbw, I, iplane_stride, IPW, NRQ, cw0, ox0, two, scale_x, support_x, plx, n_out_x, nch);
            else
                tile_hpass<5>(P, pbw, I, iplane_stride, IPW, NRQ, cw0, ox0, two, scale_x, support_x, plx, n_out_x, nch);
            PROF_MARK(4);  // H pass
            __syncthreads();  // (B) H pass done: P is free, I is complete
            PROF_MARK(5);  // barrier (B)
            if (tid == kProducerTid) {
                patch_busy = false;
                producer_advance(limit);  // the next patch streams in during this V pass
            }
            if (bg_pending) {
                wait_bg();
            }
#define B200_VPASS(NWY, NCH_) \
    tile_vpass_over<NWY, NCH_>(I, iplane_stride, IPW, NRQ, ct, rw0, oy0, tho, two, dx, dy, scale_y, support_y, ply, n_out_y)
            if (nch == 4) {
                if (nwy == 3) B200_VPASS(3, 4);
                else if (nwy == 4) B200_VPASS(4, 4);
                else B200_VPASS(5, 4);
            } else {
                if (nwy == 3) B200_VPASS(3, 3);
                else if (nwy == 4) B200_VPASS(4, 3);
                else B200_VPASS(5, 3);
            }
#undef B200_VPASS
            PROF_MARK(6);  // V pass (+ background wait, producer thread)
        } else if (kind == kCmdIdentTma) {
            // identity-size overlay: P holds the 64x32 source pixels under this tile (zero outside the overlay)
            mbar_wait(patch_full, pseq & 1u);
            ++pseq;
            if (bg_pending) {
                wait_bg();
            }
            const int xx = tid & (kTileW - 1);
            const int shift = (int)cmd.w[5];
            if (tid < kElemThreads && xx >= dx && xx < dx + two)
                for (int yy = dy + tid / kTileW; yy < dy + tho; yy += kRowSweep) {
                    uint32_t *c = ct + ct_off(yy, xx);
                    *c = over_px(*c, P[yy * kOverlayBoxW + xx + shift]);
                }
            __syncthreads();  // (B) P is free
            if (tid == kProducerTid) {
                patch_busy = false;
                producer_advance(limit);
            }
        } else {  // kCmdIdentLdg: overlay read with plain loads (source not addressable by TMA)
            if (bg_pending) {
                wait_bg();
            }
            const uint8_t *src = reinterpret_cast<const uint8_t *>((uint64_t)cmd.w[12] | ((uint64_t)cmd.w[13] << 32));
            const int64_t spitch = (int64_t)cmd.w[6];
            const int sx0 = (int)cmd.w[3], sy0 = (int)cmd.w[4];
            const int xx = tid & (kTileW - 1);
            if (tid < kElemThreads && xx < two)
                for (int yy = tid / kTileW; yy < tho; yy += kRowSweep) {
                    const uint32_t s = ld_px(src, (int64_t)(sy0 + yy) * spitch + (int64_t)(sx0 + xx) * 4);
                    uint32_t *c = ct + ct_off(dy + yy, dx + xx);
                    *c = over_px(*c, s);
                }
        }
        if (--steps_left == 0) finish_tile();
        ++pos;
    }
    if (ptid >= 0 && ptid < 4) cp_async_wait<0>();
    if (tid == kProducerTid) bulk_wait_all();  // stores still reading shared memory
#if B200COMP_PROFILE
    if (tid == B200COMP_PROFILE_TID)
        for (int i = 0; i < 10; ++i) atomicAdd(&g_prof[i], prof_acc[i]);
#endif
#undef trec
#undef c_tx0
#undef c_ty0
#undef c_tw
#undef c_th
#undef c_canvas
#undef c_out_map
}

}  // namespace b200comp
