// Device code of the B200 compositor hot path (sm_100a).  Integer-exact restatement of
// the Pillow arithmetic reached from /root/reference/compositor.py:20-21 and
// /root/reference/background_resizing.py:11-33,74-97; see DESIGN.md for the layout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200comp {

// ------------------------------------------------------------------ descriptors
struct DevCanvas {
    uint8_t *out;
    const uint8_t *bg;  // may be null -> solid
    int64_t out_pitch;
    int64_t bg_pitch;
    int64_t tile_base;  // first tile index of this canvas in the launch
    uint32_t solid;
    int32_t W, H;
    int32_t first, count;  // placement range
    int32_t tiles_x, tiles_y;
    int32_t pad_;
    const void *bg_map;   // CUtensorMap (32x32-pixel boxes, 128-byte swizzle) over bg, or null: generic loads
    const void *out_map;  // same over out, or null: generic stores
};
static_assert(sizeof(DevCanvas) == 88, "DevCanvas layout");

constexpr int kTileW = 64;
#ifndef B200COMP_TILE_H
#define B200COMP_TILE_H 64
#endif
constexpr int kTileH = B200COMP_TILE_H;  // 32 or 64 (the vertical pass walks row groups of 32).  64: half as many tile steps, so the
                                         // per-step work of a warp (decode, coefficient rows, waits) is amortised over twice the pixels,
                                         // and 9 % fewer H-pass rows (window halo); measured 4.6 % faster than 32 (profiles/r2_tile_kernel_ab.txt)
static_assert(kTileH == 32 || kTileH == 64, "tile height");
#ifndef B200COMP_CTAS_PER_SM
#define B200COMP_CTAS_PER_SM 2
#endif
constexpr int kCtasPerSm = B200COMP_CTAS_PER_SM;  // persistent CTAs resident per SM (registers and shared memory permitting)
// The tile kernel gives every compute warp a SLAB of the resident tile: kSlabW columns, all rows.  A warp runs the
// horizontal pass, the vertical pass and the over of its slab on its own -- its intermediate is private -- so
// the passes need no CTA-wide barrier.  Two more warps only move data: lane 0 of the producer warp issues every
// asynchronous load (command blocks, background tiles, source patch chunks), lane 0 of the store warp writes
// finished tiles back.
constexpr int kSlabW = 8;
constexpr int kSlabWarps = kTileW / kSlabW;
constexpr int kProducerWarp = kSlabWarps;
constexpr int kStoreWarp = kSlabWarps + 1;
constexpr int kThreads = (kSlabWarps + 2) * 32;
constexpr int kPrecisionBits = 22;
constexpr int kTileWords = kTileW * kTileH;  // one resident canvas tile: two halves of 32 pixels x kTileH rows, 128-byte swizzled
constexpr int kOverlayBoxW = kTileW + 4;     // identity overlays: box widened so its start can be 16-byte aligned
constexpr int kIdentRows = 16;               // overlay rows per chunk (one ring slot)
#ifndef B200COMP_TILE_BUFS
#define B200COMP_TILE_BUFS 2
#endif
constexpr int kTileBufs = B200COMP_TILE_BUFS;  // resident canvas tiles per CTA: one being composited, one being stored and then pre-loaded
// Source patches stream through a ring of chunks: kChunkQuads row quads (4 rows each) x 4 channel planes x the
// placement's patch width.  A chunk is one TMA box; every compute warp consumes every chunk.
#ifndef B200COMP_CHUNK_QUADS
#define B200COMP_CHUNK_QUADS 4
#endif
constexpr int kChunkQuads = B200COMP_CHUNK_QUADS;
#ifndef B200COMP_PRING
#define B200COMP_PRING 3
#endif
// Ring depth: chosen per plan at run time -- as many slots as still leave kCtasPerSm CTAs resident (the deeper the
// ring, the further the producer runs ahead of the compute warps: 4 slots measured 3.6 % faster than 3 at the headline
// workload).  kPRingMin slots are what a placement must fit with to take the fused path at all.
constexpr int kPRingMin = B200COMP_PRING;
constexpr int kPRingMax = 8;
// position in a ring of n slots whose mbarriers flip phase once per lap
struct RingPos {
    uint32_t slot = 0, phase = 0, seq = 0;
};
__device__ __forceinline__ void ring_next(RingPos &r, int n) {
    ++r.seq;
    if (++r.slot == (uint32_t)n) {
        r.slot = 0;
        r.phase ^= 1u;
    }
}
constexpr int kCmdBlk = 8;   // command records per block (one bulk copy)
constexpr int kCmdRing = 4;  // command blocks in shared memory

__host__ __device__ constexpr int coef_row_words(int nw) { return nw == 5 ? 16 : 12; }

enum : int { kStatusPatchOverflow = 1, kStatusInterOverflow = 2, kStatusStreamOverflow = 4, kStatusWatchdog = 8 };

// ------------------------------------------------------------------ command streams
// The tile kernel is persistent: CTA c of G walks the canvas tiles t = c, c + G, c + 2G, ... of a run.
// A binning pass turns (canvases, placements) into one command stream per CTA, so the tile kernel never
// chases pointers: it streams fixed-size records from a linear array.  A record is 16 words:
//   TILE      w0 kind, w1 n_steps, w2 tx0, w3 ty0, w4 tw | th << 16, w5 solid, w6 flags, w7 canvas index,
//             w8-9 bg tensor map, w10-11 out tensor map
//   RESAMPLE  w0 kind, w1 nwx | nwy << 8 | nch << 16 | NRQ << 24, w2 dx | dy << 8 | two << 16 | tho << 24,
//             w3 ox0, w4 oy0, w5 cw0 | rw0 << 16 (first patch word column / first source row quad),
//             w6 w (coefficient plane stride x), w7 h, w8 tensor map index, w9 plx word offset,
//             w10 ply word offset, w11 pwc (patch width class: word columns per chunk row),
//             w12-13 scale_x, w14-15 scale_y
//   IDENT_TMA w0 kind, w2 as above, w3 / w4 source coordinates of the tile origin's box (w3 a multiple of 4: TMA
//             boxes start on 16-byte boundaries), w5 pixels from the box start to the tile origin, w8 map index
//   IDENT_LDG w0 kind, w2 as above, w3 / w4 source coordinates of the intersection, w6 pitch, w12-13 src
struct __align__(16) Cmd {
    uint32_t w[16];
};
static_assert(sizeof(Cmd) == 64, "Cmd layout");
enum : uint32_t { kCmdTile = 0, kCmdResample = 1, kCmdIdentTma = 2, kCmdIdentLdg = 3, kCmdNop = 4, kCmdEnd = 5 };
enum : uint32_t { kTileBgTma = 1, kTileOutTma = 2, kTileHasBg = 4, kTileNoBg = 8 };

// ------------------------------------------------------------------ pixel arithmetic
// ImagingUtils.h MULDIV255
__device__ __forceinline__ uint32_t muldiv255(uint32_t a, uint32_t b) {
    uint32_t t = a * b + 128u;
    return ((t >> 8) + t) >> 8;
}
__device__ __forceinline__ uint32_t shiftfordiv255(uint32_t a) { return ((a >> 8) + a) >> 8; }

// Convert.c rgbA2rgba: RGBA -> RGBa
__device__ __forceinline__ uint32_t premultiply_px(uint32_t p) {
    const uint32_t a = p >> 24;
    if (a == 255u) return p;
    if (a == 0u) return 0u;
    const uint32_t r = muldiv255(p & 0xffu, a);
    const uint32_t g = muldiv255((p >> 8) & 0xffu, a);
    const uint32_t b = muldiv255((p >> 16) & 0xffu, a);
    return r | (g << 8) | (b << 16) | (a << 24);
}

// 255 * c / a for c, a < 256 without a division: with m = ceil(2^24 / a), (255 * c * m) >> 24 is the exact
// truncated quotient (n = 255 * c < 2^16, a < 2^8: the round-up error n * (m * a - 2^24) / (a * 2^24) stays below
// 1 / a; checked exhaustively by tests/test_host_logic.py).  One table lookup serves the three colour channels.
struct UnpremulMagic {
    uint32_t m[256];
    constexpr UnpremulMagic() : m() {
        for (uint32_t a = 1; a < 256; ++a) m[a] = ((1u << 24) + a - 1u) / a;
    }
};
__device__ const UnpremulMagic g_unpremul_magic = UnpremulMagic();

// Convert.c rgba2rgbA: RGBa -> RGBA, truncating divide, clip
__device__ __forceinline__ uint32_t unpremultiply_px(uint32_t p) {
    const uint32_t a = p >> 24;
    if (a == 255u || a == 0u) return p;
    const uint32_t m = __ldg(&g_unpremul_magic.m[a]);
    const uint32_t r = min(255u, __umulhi((p & 0xffu) * 65280u, m));  // (255 * c << 8) * m >> 32
    const uint32_t g = min(255u, __umulhi(((p >> 8) & 0xffu) * 65280u, m));
    const uint32_t b = min(255u, __umulhi(((p >> 16) & 0xffu) * 65280u, m));
    return r | (g << 8) | (b << 16) | (a << 24);
}

// AlphaComposite.c ImagingAlphaComposite, one pixel: src over dst
__device__ __forceinline__ uint32_t over_px(uint32_t d, uint32_t s) {
    const uint32_t sa = s >> 24;
    if (sa == 0u) return d;
    if (sa == 255u) return s;  // coef2 == 0 and the rounding is exact: the source pixel itself
    const uint32_t da = d >> 24;
    uint32_t coef1, outa;
    if (da == 255u) {  // opaque canvas: outa255 == 255*255, division-free
        coef1 = sa * 128u;
        outa = 255u;
    } else {
        const uint32_t blend = da * (255u - sa);
        const uint32_t outa255 = sa * 255u + blend;
        coef1 = sa * 255u * 255u * 128u / outa255;
        outa = shiftfordiv255(outa255 + 0x80u);
    }
    const uint32_t coef2 = 255u * 128u - coef1;
    const uint32_t r = shiftfordiv255((s & 0xffu) * coef1 + (d & 0xffu) * coef2 + (0x80u << 7)) >> 7;
    const uint32_t g = shiftfordiv255(((s >> 8) & 0xffu) * coef1 + ((d >> 8) & 0xffu) * coef2 + (0x80u << 7)) >> 7;
    const uint32_t b = shiftfordiv255(((s >> 16) & 0xffu) * coef1 + ((d >> 16) & 0xffu) * coef2 + (0x80u << 7)) >> 7;
    return r | (g << 8) | (b << 16) | (outa << 24);
}

__device__ __forceinline__ uint32_t over_unpremul_px(uint32_t d, uint32_t s) { return over_px(d, unpremultiply_px(s)); }

// Resample.c clip8: arithmetic shift then clamp
__device__ __forceinline__ uint32_t clip8(int32_t v) { return (uint32_t)min(255, max(0, v >> kPrecisionBits)); }

__device__ __forceinline__ uint32_t pack_clip(int32_t a0, int32_t a1, int32_t a2, int32_t a3) {
    return clip8(a0) | (clip8(a1) << 8) | (clip8(a2) << 16) | (clip8(a3) << 24);
}

__device__ __forceinline__ void mac_px(int32_t &a0, int32_t &a1, int32_t &a2, int32_t &a3, uint32_t p, int32_t k) {
    a0 += (int32_t)(p & 0xffu) * k;
    a1 += (int32_t)((p >> 8) & 0xffu) * k;
    a2 += (int32_t)((p >> 16) & 0xffu) * k;
    a3 += (int32_t)(p >> 24) * k;
}

__device__ __forceinline__ uint32_t ld_px(const uint8_t *base, int64_t off) {
    return __ldg(reinterpret_cast<const uint32_t *>(base + off));
}

}  // namespace b200comp
