// Fused resample + alpha-over tile kernel (sm_100a), the hot kernel of the path.
//
// One CTA owns one 64x32 tile of one output canvas.  The tile lives in shared memory while the
// CTA walks the canvas' placements in z-order (compositor.py:12-21).  For every placement that
// touches the tile:
//   1. stage   the needed source patch: 128-bit loads of RGBA pixels, premultiply (Convert.c
//              rgbA2rgba), byte-transpose into four channel planes (4 consecutive pixels of one
//              channel per 32-bit word)
//   2. H pass  lane <-> output column; Pillow's 22-bit fixed-point taps are held as three byte
//              planes (k = b0 + 256*b1 + 65536*b2, b2 signed) so four taps cost three dp4a and no
//              byte unpacking; result rounded + clipped to uint8 (ImagingResampleHorizontal_8bpc)
//              and written transposed (4 consecutive ROWS of one channel per word)
//   3. V pass  lane <-> output row, same dp4a scheme (ImagingResampleVertical_8bpc), then
//              un-premultiply (rgba2rgbA) and alpha-over (AlphaComposite.c) onto the resident tile
// The tile is written to HBM once.  Arithmetic is integer-exact: dp4a partial sums wrap modulo
// 2^32 and the true accumulator fits in int32, exactly as Pillow's int accumulator.
#pragma once
#include "kernels.cuh"

namespace b200comp {

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
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
    uint32_t *dst;    // [sh][4][w4p]
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
        uint32_t *o = d.dst + ((int64_t)r * 4) * d.w4p + g;
        if (gx >= d.sw) {  // padding words
            o[0] = 0u; o[d.w4p] = 0u; o[2 * d.w4p] = 0u; o[3 * d.w4p] = 0u;
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
        if (((A ^ (A >> 1)) & 0x7f7f7f7fu) == 0u) {
            // every alpha is 0 or 255: MULDIV255(c, a) is c or 0 -> mask the colours with the alpha bytes
            R &= A; G &= A; B &= A;
        } else {
            transpose4(premultiply_px(p0), premultiply_px(p1), premultiply_px(p2), premultiply_px(p3), R, G, B, A);
        }
        o[0] = R; o[d.w4p] = G; o[2 * d.w4p] = B; o[3 * d.w4p] = A;
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

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// pull the coefficient rows this lane will need into L1 while the patch is being staged
__device__ __forceinline__ void prefetch_coeffs(const uint32_t *__restrict__ pl, int nw, int n_out, int idx) {
    for (int q = 0; q < 3 * nw; ++q) prefetch_l1(pl + (int64_t)q * n_out + idx);
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncg = (two + 31) >> 5;  // 1 or 2 column groups of 32
    const int cg = warp % ncg;
    const int rstep = kWarps / ncg;
    const int jj = cg * 32 + lane;
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
    for (int rq = warp / ncg; rq < NRQ; rq += rstep) {
        uint32_t o[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const uint32_t *row = P + (rq * 4 + rr) * PBW + wbase;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c == 3 && nch == 3) break;
                uint32_t a0 = 1u << (kPrecisionBits - 1), a1 = 0u;  // rounding term rides in the low plane
                int32_t a2 = 0;
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    const uint32_t wd = row[c * (PBW >> 2) + i];
                    a0 = dp4a_uu(wd, k0[i], a0);
                    a1 = dp4a_uu(wd, k1[i], a1);
                    a2 = dp4a_us(wd, k2[i], a2);
                }
                const uint32_t v = clip8i((int32_t)(a0 + (a1 << 8) + ((uint32_t)a2 << 16)));
                if (rr == 0) o[c] = v;
                else if (rr == 1) o[c] = __byte_perm(o[c], v, 0x3240);
                else if (rr == 2) o[c] = __byte_perm(o[c], v, 0x3410);
                else o[c] = __byte_perm(o[c], v, 0x4210);
            }
        }
        uint32_t *d = I + jj * IPW + rq;
        d[0] = o[0];
        d[iplane_stride] = o[1];
        d[2 * iplane_stride] = o[2];
        if (nch == 4) d[3 * iplane_stride] = o[3];
    }
}

// ---- V pass + un-premultiply + over -------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void tile_vpass_over(const uint32_t *__restrict__ I, int iplane_stride, int IPW,
                                                uint32_t *__restrict__ ctile, int rw0, int oy0, int tho, int two,
                                                int tile_dx, int tile_dy, double scale, double support,
                                                const uint32_t *__restrict__ ply, int n_out, int nch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= tho) return;
    const int y = oy0 + lane;
    const int wbase = (first_tap(y, scale, support) >> 2) - rw0;
    uint32_t k0[NW], k1[NW], k2[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        k0[i] = __ldg(ply + (int64_t)(0 * NW + i) * n_out + y);
        k1[i] = __ldg(ply + (int64_t)(1 * NW + i) * n_out + y);
        k2[i] = __ldg(ply + (int64_t)(2 * NW + i) * n_out + y);
    }
    uint32_t *crow = ctile + (tile_dy + lane) * kCtPitch + tile_dx;
    for (int x = warp; x < two; x += kWarps) {
        const uint32_t *col = I + x * IPW + wbase;
        int32_t acc[4];
        acc[3] = 255 << kPrecisionBits;  // opaque patch: alpha is exactly 255
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c == 3 && nch == 3) break;
            uint32_t a0 = 1u << (kPrecisionBits - 1), a1 = 0u;
            int32_t a2 = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const uint32_t wd = col[c * iplane_stride + i];
                a0 = dp4a_uu(wd, k0[i], a0);
                a1 = dp4a_uu(wd, k1[i], a1);
                a2 = dp4a_us(wd, k2[i], a2);
            }
            acc[c] = (int32_t)(a0 + (a1 << 8) + ((uint32_t)a2 << 16));
        }
        // Alpha tests on the raw accumulator (clip8(acc) == 0 / == 255).  Do NOT test the clamped value:
        // CUDA 12.9 ptxas folds `clamp(x) == 255` into VIMNMX.RELU's predicate output with the wrong
        // sense on sm_100a (partially transparent pixels took the opaque branch).
        if (acc[3] < (1 << kPrecisionBits)) continue;  // transparent: canvas pixel unchanged
        const bool opaque = acc[3] >= (255 << kPrecisionBits);
        const uint32_t s = clip8i(acc[0]) | (clip8i(acc[1]) << 8) | (clip8i(acc[2]) << 16) | (clip8i(acc[3]) << 24);
        crow[x] = opaque ? s : over_px(crow[x], unpremultiply_px(s));
    }
}

struct DevPlacementT {
    const uint8_t *src;    // mode 0: w x h overlay composited as is (raw RGBA)
    const uint32_t *plx;   // [3*nwx][w] coefficient byte planes of the horizontal pass
    const uint32_t *ply;   // [3*nwy][h] vertical pass
    const void *tmap;      // mode 1: CUtensorMap over the prepared cutout, box = (pbw/4 words, 4 planes, nrbox rows)
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

constexpr int kDescCache = 64;  // placement descriptors cached in shared memory per CTA
constexpr int kDescWords = sizeof(DevPlacementT) / 4;

// grid = (max tiles per canvas, n canvases): blockIdx.y is the canvas, blockIdx.x its tile.
__global__ void __launch_bounds__(kThreads, 2)
composite_tiles_kernel(const DevCanvas *__restrict__ canvases, const DevPlacementT *__restrict__ placements,
                       int patch_words, int inter_words, int *__restrict__ status) {
    extern __shared__ __align__(128) uint32_t smem[];
    uint32_t *ctile = smem;                   // kTileH * kCtPitch (8320 B: keeps P 128-byte aligned for TMA)
    uint32_t *P = ctile + kTileH * kCtPitch;  // patch_words
    uint32_t *I = P + patch_words;            // inter_words
    __shared__ __align__(16) uint32_t desc_words[kDescCache * kDescWords];
    __shared__ uint32_t hit_mask[kDescCache / 32];
    __shared__ uint8_t hit_list[kDescCache];
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ uint32_t alpha_bits[2];  // OR of the alpha summary over the current patch (double buffered)

    const DevCanvas cv = canvases[blockIdx.y];
    const int local = blockIdx.x;
    if (local >= cv.tiles_x * cv.tiles_y) return;  // canvases of different sizes share one grid
    const int ty = local / cv.tiles_x, tx = local - ty * cv.tiles_x;
    const int tx0 = tx * kTileW, ty0 = ty * kTileH;
    const int tx1 = min(cv.W, tx0 + kTileW), ty1 = min(cv.H, ty0 + kTileH);
    const int tw = tx1 - tx0, th = ty1 - ty0;

    // ---- placement descriptors -> shared memory (one coalesced pass), canvas tile -> shared memory ----
    {
        const int n0 = min(cv.count, kDescCache);
        const uint32_t *g = reinterpret_cast<const uint32_t *>(placements + cv.first);
        for (int i = threadIdx.x; i < n0 * kDescWords; i += kThreads) desc_words[i] = __ldg(g + i);
    }
    // the background tile streams in asynchronously (cp.async); it is first needed by an over step
    {
        const int xx = threadIdx.x & (kTileW - 1), y0 = threadIdx.x / kTileW;  // 4 rows per sweep
        const uint8_t *g = cv.bg + (int64_t)(ty0 + y0) * cv.bg_pitch + (int64_t)(tx0 + xx) * 4;
        uint32_t *d = ctile + y0 * kCtPitch + xx;
        const int64_t gstep = (int64_t)(kThreads / kTileW) * cv.bg_pitch;
        if (xx < tw) {
#pragma unroll
            for (int k = 0; k < (kTileH + kThreads / kTileW - 1) / (kThreads / kTileW); ++k) {
                if (y0 + k * (kThreads / kTileW) < th) {
                    if (cv.bg) cp_async4(d, g); else *d = cv.solid;
                }
                g += gstep;
                d += (kThreads / kTileW) * kCtPitch;
            }
        }
    }
    __syncthreads();
    const DevPlacementT *desc = reinterpret_cast<const DevPlacementT *>(desc_words);
    if (threadIdx.x == 0) {
        mbar_init(&tma_bar, 1);
        alpha_bits[0] = 0u;
        alpha_bits[1] = 0u;
    }
    uint32_t n_res = 0;  // resampled placements seen so far (selects the alpha_bits slot)
    uint32_t tma_phase = 0;  // parity of the next TMA completion to wait for
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // Geometry of a resampled placement on this tile; all threads compute it (cheap, no memory).
    struct Geo {
        int ix0, iy0, two, tho, ox0, oy0, cw0, rw0, NRQ, bq0, bq1;
    };
    auto geometry = [&](const DevPlacementT &d) {
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
        g.bq0 = w_first >> 2;                   // alpha summary blocks (4 words) the patch touches
        g.bq1 = min(w_last >> 2, d.wq - 1);
        g.cw0 = w_first & ~3;  // TMA boxes start on 16-byte boundaries
        g.rw0 = first_tap(g.oy0, d.scale_y, d.support_y) >> 2;
        g.NRQ = (first_tap(iy1 - 1 - d.y, d.scale_y, d.support_y) >> 2) + d.nwy - g.rw0;
        return g;
    };
    // one elected thread starts the TMA of placement `d`'s source patch into P
    auto issue_patch = [&](const DevPlacementT &d) {
        const Geo g = geometry(d);
        mbar_expect_tx(&tma_bar, (uint32_t)d.pbw * (uint32_t)d.nrbox * 4u);
        tma_load_patch(P, d.tmap, g.cw0, 4 * g.rw0, &tma_bar);
    };

    // ---- z-order walk, kDescCache placements at a time ----
    for (int group = 0; group < cv.count; group += kDescCache) {
        const int n_cached = min(cv.count - group, kDescCache);
        if (group > 0) {
            __syncthreads();  // everyone is done with the previous group's descriptors
            const uint32_t *g = reinterpret_cast<const uint32_t *>(placements + cv.first + group);
            for (int i = threadIdx.x; i < n_cached * kDescWords; i += kThreads) desc_words[i] = __ldg(g + i);
            __syncthreads();
        }
        // which cached placements touch this tile -> ordered hit list
        if (threadIdx.x < kDescCache) {
            bool hit = false;
            if ((int)threadIdx.x < n_cached) {
                const DevPlacementT &d = desc[threadIdx.x];
                hit = max(tx0, d.x) < min(tx1, d.x + d.w) && max(ty0, d.y) < min(ty1, d.y + d.h);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) hit_mask[warp] = m;
        }
        __syncthreads();
        if (threadIdx.x < kDescCache) {
            const uint32_t m0 = hit_mask[0], m1 = hit_mask[1];
            const uint32_t mine = warp == 0 ? m0 : m1;
            if ((mine >> lane) & 1u)
                hit_list[(warp == 0 ? 0 : __popc(m0)) + __popc(mine & ((1u << lane) - 1u))] = (uint8_t)threadIdx.x;
        }
        const int n_hits = __popc(hit_mask[0]) + __popc(hit_mask[1]);
        __syncthreads();
        // first resampled placement of the group: start its TMA now
        int next_res = 0;
        while (next_res < n_hits && desc[hit_list[next_res]].mode == 0) ++next_res;
        if (next_res < n_hits && threadIdx.x == 0) issue_patch(desc[hit_list[next_res]]);

        for (int k = 0; k < n_hits; ++k) {
            const DevPlacementT &d = desc[hit_list[k]];
            if (d.mode == 0) {
                // identity-size placement: plain over straight from the cutout
                const int ix0 = max(tx0, d.x), iy0 = max(ty0, d.y);
                const int two = min(tx1, d.x + d.w) - ix0, tho = min(ty1, d.y + d.h) - iy0;
                cp_async_wait_all();
                __syncthreads();
                const int xx = threadIdx.x & (kTileW - 1);
                if (xx < two) {
                    for (int yy = threadIdx.x / kTileW; yy < tho; yy += kThreads / kTileW) {
                        const uint32_t s = ld_px(d.src, (int64_t)(iy0 + yy - d.y) * d.src_pitch + (int64_t)(ix0 + xx - d.x) * 4);
                        uint32_t *c = ctile + (iy0 + yy - ty0) * kCtPitch + (ix0 + xx - tx0);
                        *c = over_px(*c, s);
                    }
                }
                __syncthreads();
                continue;
            }
            const Geo g = geometry(d);
            const int IPW = g.NRQ | 1;
            const int iplane_stride = kTileW * IPW;
            const bool fits = d.pbw * d.nrbox <= patch_words && 4 * iplane_stride <= inter_words && 4 * g.NRQ <= d.nrbox;
            if (!fits && threadIdx.x == 0) atomicOr(status, kStatusPatchOverflow);  // host sizing bug: flagged
            // alpha summary of the patch (global loads overlap the TMA already in flight)
            {
                const int nbw = g.bq1 - g.bq0 + 1;
                const int rq1 = min(g.rw0 + g.NRQ, d.sh4);
                const int nb = nbw * (rq1 - g.rw0);
                uint32_t bits = 0u;
                for (int i = threadIdx.x; i < nb; i += kThreads) {
                    const int br = i / nbw, bc = i - br * nbw;
                    bits |= __ldg(d.flags + (int64_t)(g.rw0 + br) * d.wq + g.bq0 + bc);
                }
                bits = __reduce_or_sync(0xffffffffu, bits);
                if (lane == 0 && bits) atomicOr(&alpha_bits[n_res & 1u], bits);
                if (threadIdx.x == 0) alpha_bits[(n_res + 1u) & 1u] = 0u;  // slot of the next resampled placement
            }
            // coefficient rows -> L1 while the patch is in flight
            if (warp < 2 && warp * 32 + lane < g.two) prefetch_coeffs(d.plx, d.nwx, d.w, g.ox0 + warp * 32 + lane);
            if (warp == 2 && lane < g.tho) prefetch_coeffs(d.ply, d.nwy, d.h, g.oy0 + lane);
            mbar_wait(&tma_bar, tma_phase);  // source patch has landed in P
            tma_phase ^= 1u;
            __syncthreads();  // alpha_bits complete; everyone has consumed this TMA phase
            const uint32_t abits = alpha_bits[n_res & 1u];
            ++n_res;
            const bool transparent = (abits & 1u) == 0u;  // nothing but alpha 0: the canvas does not change
            const int nch = (abits & 2u) ? 4 : 3;         // every alpha 255: skip the alpha plane
            if (transparent) {
                next_res = k + 1;
                while (next_res < n_hits && desc[hit_list[next_res]].mode == 0) ++next_res;
                if (next_res < n_hits && threadIdx.x == 0) issue_patch(desc[hit_list[next_res]]);
                __syncthreads();  // everyone has read alpha_bits before its slot is recycled
                continue;
            }
            if (fits) {
                if (d.nwx == 3)
                    tile_hpass<3>(P, d.pbw, I, iplane_stride, IPW, g.NRQ, g.cw0, g.ox0, g.two, d.scale_x, d.support_x, d.plx, d.w, nch);
                else if (d.nwx == 4)
                    tile_hpass<4>(P, d.pbw, I, iplane_stride, IPW, g.NRQ, g.cw0, g.ox0, g.two, d.scale_x, d.support_x, d.plx, d.w, nch);
                else
                    tile_hpass<5>(P, d.pbw, I, iplane_stride, IPW, g.NRQ, g.cw0, g.ox0, g.two, d.scale_x, d.support_x, d.plx, d.w, nch);
            }
            cp_async_wait_all();  // background tile (no-op after the first over)
            __syncthreads();      // H pass done: P is free, I is complete
            // next resampled placement: its patch streams in while this one runs its V pass
            next_res = k + 1;
            while (next_res < n_hits && desc[hit_list[next_res]].mode == 0) ++next_res;
            if (next_res < n_hits && threadIdx.x == 0) issue_patch(desc[hit_list[next_res]]);
            if (fits) {
                if (d.nwy == 3)
                    tile_vpass_over<3>(I, iplane_stride, IPW, ctile, g.rw0, g.oy0, g.tho, g.two, g.ix0 - tx0, g.iy0 - ty0, d.scale_y, d.support_y, d.ply, d.h, nch);
                else if (d.nwy == 4)
                    tile_vpass_over<4>(I, iplane_stride, IPW, ctile, g.rw0, g.oy0, g.tho, g.two, g.ix0 - tx0, g.iy0 - ty0, d.scale_y, d.support_y, d.ply, d.h, nch);
                else
                    tile_vpass_over<5>(I, iplane_stride, IPW, ctile, g.rw0, g.oy0, g.tho, g.two, g.ix0 - tx0, g.iy0 - ty0, d.scale_y, d.support_y, d.ply, d.h, nch);
            }
            __syncthreads();
        }
    }

    // ---- write the tile once ----
    cp_async_wait_all();
    __syncthreads();
    if (tw == kTileW && ((reinterpret_cast<uintptr_t>(cv.out) | (uintptr_t)cv.out_pitch) & 15u) == 0) {
        // full-width tile, 16-byte aligned rows: one 128-bit store per 4 pixels
        constexpr int kRows = kThreads / 16;  // rows per sweep
        const int x4 = (threadIdx.x & 15) * 4, y0 = threadIdx.x >> 4;
        uint8_t *g = cv.out + (int64_t)(ty0 + y0) * cv.out_pitch + (int64_t)(tx0 + x4) * 4;
        const uint32_t *c = ctile + y0 * kCtPitch + x4;
#pragma unroll
        for (int k = 0; k < (kTileH + kRows - 1) / kRows; ++k) {
            if (y0 + kRows * k < th)
                *reinterpret_cast<uint4 *>(g) = make_uint4(c[0], c[1], c[2], c[3]);
            g += (int64_t)kRows * cv.out_pitch;
            c += kRows * kCtPitch;
        }
    } else {
        const int xx = threadIdx.x & (kTileW - 1);
        if (xx < tw)
            for (int yy = threadIdx.x / kTileW; yy < th; yy += kThreads / kTileW)
                *reinterpret_cast<uint32_t *>(cv.out + (int64_t)(ty0 + yy) * cv.out_pitch + (int64_t)(tx0 + xx) * 4) =
                    ctile[yy * kCtPitch + xx];
    }
}

}  // namespace b200comp
