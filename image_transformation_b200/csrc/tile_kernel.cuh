// Fused resample + alpha-over tile kernel (sm_100a), the hot kernel of the path, and the binning pass
// that feeds it.
//
// The kernel is persistent: CTA c walks the 64 x kTileH canvas tiles of command stream c (written by the binning
// kernels below).  A tile lives in shared memory while the placements that touch it are composited in z-order
// (compositor.py:12-21); the binning pass has already dropped the placements that cannot show (transparent
// source patch, or hidden by a later opaque placement that covers the tile).  Every compute warp owns a slab of
// kSlabW tile columns and does, per resampled placement and on its own:
//   1. H pass  over the source patch, which streams through a ring of TMA chunks (kChunkQuads row quads x 4
//              channel planes of the PREPARED cutout: premultiplied (Convert.c rgbA2rgba), channel-planar, word
//              (rq, c, g, r) = pixels 4g..4g+3 of channel c in source row 4*rq + r, so one 128-bit shared-memory
//              load feeds four rows).  Pillow's 22-bit fixed-point taps are held as three byte planes
//              (k = b0 + 256*b1 + 65536*b2, b2 signed): four taps cost three dp4a and no byte unpacking.  The
//              result is rounded + clipped to uint8 (ImagingResampleHorizontal_8bpc) and written transposed
//              (4 consecutive ROWS of one channel per word) into the warp's private intermediate.
//   2. V pass  lane <-> output row, same dp4a scheme (ImagingResampleVertical_8bpc), one 128-bit load per four
//              channels, then un-premultiply (rgba2rgbA) and alpha-over (AlphaComposite.c) onto the resident tile.
// Nothing in a step synchronises the CTA: warps meet only on mbarriers (chunk landed / chunk released, tile
// ready / tile done).  The tile is written to HBM once (TMA store).  Arithmetic is integer-exact: dp4a partial
// sums wrap modulo 2^32 and the true accumulator fits in int32, exactly as Pillow's int accumulator.
#pragma once
#include "tile_common.cuh"

namespace b200comp {

// ---- prepare: RGBA cutout -> premultiplied, channel-planar, row-quad interleaved words --------------
// One launch per plan run converts every distinct cutout the tile kernel resamples into the layout its dp4a
// passes consume: dst[rq][c][g][r] (32-bit words) = pixels 4g..4g+3 of channel c in row 4*rq + r, colours
// premultiplied (Convert.c rgbA2rgba).  The four rows of a word column are 16 contiguous bytes, so (a) any word
// column is a legal TMA box start and (b) the horizontal pass reads four rows with one LDS.128.  Pixels past
// the row end and rows past the last one are zero.
struct PrepDesc {
    const uint8_t *src;
    uint32_t *dst;    // [sh4][4][w4][4]; null: only the alpha summary is produced (overlays composited as they are)
    uint32_t *flags;  // [sh4][wq] alpha summary of each 4-row x 16-pixel block (zeroed before the launch):
                      // bit 0 = some alpha != 0, bit 1 = some alpha != 255
    int64_t src_pitch;
    int32_t sw, sh;
    int32_t w4;       // word columns per row (ceil(sw / 4))
    int32_t wq;       // summary blocks per row (ceil(w4 / 4))
    int32_t vec_ok;   // src 16-byte aligned with pitch % 16 == 0
    int32_t pad_;
};

__global__ void __launch_bounds__(256) prepare_cutouts_kernel(const PrepDesc *__restrict__ descs) {
    const PrepDesc d = descs[blockIdx.y];
    const int sh4 = (d.sh + 3) >> 2;
    const int64_t total = (int64_t)sh4 * d.w4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int rq = (int)(i / d.w4), g = (int)(i - (int64_t)rq * d.w4);
        const int gx = 4 * g;
        const int nv = min(4, d.sw - gx);
        uint32_t R[4], G[4], B[4], A[4];
        uint32_t bits = 0u;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int y = 4 * rq + rr;
            R[rr] = G[rr] = B[rr] = A[rr] = 0u;
            if (y >= d.sh) continue;
            const uint8_t *rowp = d.src + (int64_t)y * d.src_pitch + (int64_t)gx * 4;
            uint32_t p0, p1 = 0u, p2 = 0u, p3 = 0u;
            if (d.vec_ok && nv == 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rowp));
                p0 = v.x; p1 = v.y; p2 = v.z; p3 = v.w;
            } else {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(rowp);
                p0 = __ldg(q);
                if (nv > 1) p1 = __ldg(q + 1);
                if (nv > 2) p2 = __ldg(q + 2);
                if (nv > 3) p3 = __ldg(q + 3);
            }
            uint32_t r, gg, b, a;
            transpose4(p0, p1, p2, p3, r, gg, b, a);
            // alpha summary: lets the tile kernel skip fully transparent patches and the alpha plane of
            // fully opaque ones (pixels past the row end count as neither)
            const uint32_t a_all = nv < 4 ? (a | (0xffffffffu << (8 * nv))) : a;
            bits |= (a != 0u ? 1u : 0u) | (a_all != 0xffffffffu ? 2u : 0u);
            if (d.dst) {
                if (((a ^ (a >> 1)) & 0x7f7f7f7fu) == 0u) {
                    // every alpha is 0 or 255: MULDIV255(c, a) is c or 0 -> mask the colours with the alpha bytes
                    r &= a; gg &= a; b &= a;
                } else {
                    transpose4(premultiply_px(p0), premultiply_px(p1), premultiply_px(p2), premultiply_px(p3), r, gg, b, a);
                }
                R[rr] = r; G[rr] = gg; B[rr] = b; A[rr] = a;
            }
        }
        if (d.dst) {
            uint4 *o = reinterpret_cast<uint4 *>(d.dst + ((int64_t)rq * 4 * d.w4 + g) * 4);
            o[0] = make_uint4(R[0], R[1], R[2], R[3]);
            o[d.w4] = make_uint4(G[0], G[1], G[2], G[3]);
            o[2 * d.w4] = make_uint4(B[0], B[1], B[2], B[3]);
            o[3 * d.w4] = make_uint4(A[0], A[1], A[2], A[3]);
        }
        if (bits) atomicOr(d.flags + (int64_t)rq * d.wq + (g >> 2), bits);
    }
}

struct DevPlacementT {
    const uint8_t *src;    // mode 0: w x h overlay composited as is (raw RGBA)
    const uint32_t *plx;   // [3*nwx][w] coefficient byte planes of the horizontal pass
    const uint32_t *ply;   // [3*nwy][h] vertical pass
    const void *tmap;      // mode 1: CUtensorMap over the prepared cutout, box = (4 * pwc words, 4 planes, kChunkQuads row quads)
                           // mode 0: CUtensorMap over the raw overlay, box = 68 x 16 pixels (null: generic loads)
    const uint32_t *flags; // mode 1: alpha summary of the prepared cutout, [sh4][wq] (see PrepDesc)
    double scale_x, support_x;  // sw / w and 3 * max(1, scale): exactly the host builder's doubles
    double scale_y, support_y;
    int32_t src_pitch;     // bytes (mode 0)
    int32_t sw, sh;
    int32_t x, y, w, h;    // destination box
    int32_t nwx, nwy;      // words per output sample (3, 4 or 5)
    int32_t mode;          // 0 = plain over, 1 = resample in the tile kernel
    int32_t pwc;           // patch width class: word columns (4 pixels of one channel) per chunk row
    int32_t replace;       // mode 1: store the resampled pixel instead of compositing it (stand-alone resize)
    int32_t wq, sh4;       // alpha summary extent: blocks per row, block rows
};
static_assert(sizeof(DevPlacementT) == 128, "DevPlacementT layout");

// Geometry of a resampled placement on a tile (identical doubles to the host table builder).
struct Geo {
    int ix0, iy0, two, tho, ox0, oy0, cw0, pw, rw0, NRQ, bq0, bq1;
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
    g.cw0 = w_first;  // a word column of the prepared layout is 16 bytes (4 rows): any column is a legal TMA start
    g.pw = w_last - w_first + 1;
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
                 int mask_chunks, int iw_words, unsigned long long *__restrict__ cursor,
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
            const bool fits = g.pw <= d.pwc && 4 * kSlabW * (g.NRQ | 1) <= iw_words && g.NRQ <= 255 &&
                              g.cw0 < 65536 && g.rw0 < 65536;
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
                int64_t *__restrict__ stream_len, unsigned long long *__restrict__ cursor, Cmd *__restrict__ streams, int64_t capacity,
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
        stream_len[c] = len;
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
            // buffers or tile widths the 16-byte path cannot take: 4-byte accesses, lane = column (two per lane), eight
            // rows of loads in flight before their stores (one load-store pair at a time costs a DRAM round trip each)
            const uint8_t *bg = cv.bg;
            uint8_t *out = cv.out;
            const int64_t bgp = cv.bg_pitch, outp = cv.out_pitch;
            const uint32_t solid = cv.solid;
            for (int y0 = 0; y0 < th; y0 += 8) {
                uint32_t v[8][2];
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int yy = y0 + k, xx = lane + 32 * h;
                        v[k][h] = (bg && yy < th && xx < tw) ? ld_px(bg, (int64_t)(ty0 + yy) * bgp + (int64_t)(tx0 + xx) * 4) : solid;
                    }
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int yy = y0 + k, xx = lane + 32 * h;
                        if (yy < th && xx < tw)
                            *reinterpret_cast<uint32_t *>(out + (int64_t)(ty0 + yy) * outp + (int64_t)(tx0 + xx) * 4) = v[k][h];
                    }
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
                w[1] = (uint32_t)d.nwx | ((uint32_t)d.nwy << 8) | (((om >> lane) & 1u) ? (3u << 16) : (4u << 16)) |
                       (d.replace ? (0x80u << 16) : 0u) | ((uint32_t)g.NRQ << 24);
                w[2] = (uint32_t)(g.ix0 - tx0) | ((uint32_t)(g.iy0 - ty0) << 8) | ((uint32_t)g.two << 16) | ((uint32_t)g.tho << 24);
                w[3] = (uint32_t)g.ox0;
                w[4] = (uint32_t)g.oy0;
                w[5] = (uint32_t)g.cw0 | ((uint32_t)g.rw0 << 16);
                w[6] = (uint32_t)d.w;
                w[7] = (uint32_t)d.h;
                w[8] = (uint32_t)((reinterpret_cast<const uint8_t *>(d.tmap) - maps_base) >> 7);
                w[9] = (uint32_t)(d.plx - tables_base);
                w[10] = (uint32_t)(d.ply - tables_base);
                w[11] = (uint32_t)d.pwc;
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

#ifndef B200COMP_L2_PREFETCH
#define B200COMP_L2_PREFETCH 0  // measured +-0 (profiles/r2_tile_kernel_ab.txt): the patch loads are not what the warps wait for
#endif
// ---- the persistent tile kernel ------------------------------------------------------------------
// Shared-memory rendezvous points of one CTA.  Every ring uses the n-th use of a slot <-> phase parity
// (n / ring size) & 1 convention; a producer re-fills a slot only after the matching `empty` / `free` phase.
struct SlabBars {
    uint64_t p_full[kPRingMax], p_empty[kPRingMax];                      // patch chunks (n_ring of them in use)
    uint64_t t_ready[kTileBufs], t_done[kTileBufs], t_free[kTileBufs];  // resident tiles
    uint64_t c_full[kCmdRing], c_empty[kCmdRing];                        // command blocks
};

// A wait that does not come true within ~1.5 s is a protocol bug: record who waited for what in the plan's debug
// words, raise kStatusWatchdog (every other wait then gives up as well) and leave the kernel, so a test run
// reports the fault instead of hanging the GPU.
constexpr int kProducerSleepNs = 100;
constexpr int kStoreSleepNs = 2000;
struct Watch {
    int *status;
    uint32_t *dbg;
};
__device__ __noinline__ void watchdog_fire(const Watch &w, uint32_t tag, uint32_t parity, uint32_t aux) {
    if ((atomicOr(w.status, kStatusWatchdog) & kStatusWatchdog) == 0) {
        w.dbg[0] = tag;
        w.dbg[1] = blockIdx.x;
        w.dbg[2] = threadIdx.x;
        w.dbg[3] = parity;
        w.dbg[4] = aux;
        __threadfence();
    }
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, const Watch &w, uint32_t tag, uint32_t aux) {
    if (mbar_try(bar, parity)) return;
    uint32_t spins = 0;
    for (;;) {
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
        if (mbar_try(bar, parity)) return;
        if (++spins < (SLEEP_NS > 0 ? (1u << 22) : (1u << 26))) continue;
        if ((*reinterpret_cast<volatile int *>(w.status) & kStatusWatchdog) == 0) watchdog_fire(w, tag, parity, aux);
        asm volatile("exit;");
    }
}
enum : uint32_t {
    kTagCmdFull = 1, kTagCmdEmpty, kTagTileReady, kTagTileDone, kTagTileFree, kTagPatchFull, kTagPatchEmpty,
    kTagRoleConsumer = 0x100, kTagRoleProducer = 0x200, kTagRoleStore = 0x300
};

// ---- H pass of one warp's slab --------------------------------------------------------------------
// Lane = (column jl = lane & 7 of the slab, item lane q = lane >> 3).  The work items of a chunk are its
// (row quad, channel) pairs, nch per quad; the four lanes of a column take items q, q + 4, ... -- with four
// channels a lane keeps one channel, with three (opaque patch: the alpha plane is 255 after both passes, because
// |sum(k) - 2^22| <= taps, and is not computed) the twelve items of a full chunk still split evenly.  One item =
// four source rows of one channel: NW LDS.128, 12 * NW dp4a, one word of four clipped bytes stored to
// Iw[jl][rq][c].  An item whose source words are all zero -- transparent pixels: premultiplied colours are zero
// as well -- stores zero without computing (whole-warp vote, so the pipes see no divergence).
// B200COMP_ABLATE: MEASUREMENT builds only (tools/ablate.sh) -- the pixels are wrong.  Parts of a step are compiled out to
// see what they cost inside the running pipeline: bit 0 the H pass arithmetic, bit 1 the V pass, bit 2 the V pass's taps /
// over / store (its loads, votes and addressing stay), bit 3 the V pass's window start and coefficient row, bit 4 the patch
// copies (slots are declared full without a TMA load).
#ifndef B200COMP_ABLATE
#define B200COMP_ABLATE 0
#endif

template <int NW, int NCH>
__device__ __forceinline__ void slab_hpass(const uint32_t *__restrict__ P, int slot_words, int n_ring, SlabBars *bars, RingPos &rp,
                                           const Watch &watch, uint32_t *__restrict__ Iw, int CS, int NRQ, int pwc,
                                           int cw0, int j, bool active, double scale, double support,
                                           const uint32_t *__restrict__ plx, int n_out) {
    const int lane = threadIdx.x & 31, jl = lane & 7, q = lane >> 3;
    const int wb4 = ((first_tap(j, scale, support) >> 2) - cw0) * 4;
    uint32_t k0[NW], k1[NW], k2[NW];
    load_coef_row<NW>(plx, j, k0, k1, k2);
    (void)n_out;
    uint32_t QSb = 64u * (uint32_t)pwc, PSb = 16u * (uint32_t)pwc;
    asm volatile("" : "+r"(QSb), "+r"(PSb));
    const int ch0 = (NCH == 4 || q < 3) ? q : 0, rq0 = (NCH == 4 || q < 3) ? 0 : 1;
    const uint32_t lane_src = smem_u32(P) + 4u * (uint32_t)wb4;
    const uint32_t lane_dst = smem_u32(Iw) + 4u * (uint32_t)(jl * CS + ch0);
    const uint32_t soff0 = active ? (uint32_t)rq0 * QSb + (uint32_t)ch0 * PSb : 0u;
#pragma unroll 1
    for (int q0 = 0; q0 < NRQ; q0 += kChunkQuads, ring_next(rp, n_ring)) {
        const int s = (int)rp.slot;
        mbar_wait(&bars->p_full[s], rp.phase, watch, kTagRoleConsumer | kTagPatchFull, rp.seq);
        const uint32_t slot = lane_src + (uint32_t)(s * slot_words) * 4u;
        const int n_items = NCH * min(kChunkQuads, NRQ - q0);
        const int n_iter = (n_items + 3) >> 2;
        int it = q, ch = ch0;
        uint32_t src = slot + soff0;
        uint32_t dst = lane_dst + 16u * (uint32_t)(q0 + rq0);
#pragma unroll 1
        for (int t = 0; t < n_iter; ++t) {
            const bool valid = active && it < n_items;
            uint4 v[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) v[i] = lds128((valid ? src : slot) + 16u * i);
            uint32_t o = 0u;
            {
                uint32_t wd[NW];
                int32_t sv[4];
#pragma unroll
                for (int i = 0; i < NW; ++i) wd[i] = v[i].x;
                sv[0] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
                for (int i = 0; i < NW; ++i) wd[i] = v[i].y;
                sv[1] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
                for (int i = 0; i < NW; ++i) wd[i] = v[i].z;
                sv[2] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
                for (int i = 0; i < NW; ++i) wd[i] = v[i].w;
                sv[3] = tap_sum<NW>(wd, k0, k1, k2);
                o = pack2_clip(sv[0], sv[1], pack2_clip(sv[2], sv[3], 0u));
            }
            if (valid) sts32(dst, o);
            it += 4;
            if (NCH == 4) {
                src += QSb;
                dst += 16u;
            } else {
                ch += 1;
                src += QSb + PSb;
                dst += 20u;
                if (ch >= 3) {
                    ch -= 3;
                    src += QSb - 3u * PSb;
                    dst += 4u;
                }
            }
        }
        named_bar_arrive(1 + s, (kSlabWarps + 1) * 32);
    }
}

template <int NW>
__device__ __forceinline__ void vcol3(const uint4 (&v)[NW], uint32_t cpx, const uint32_t (&k0)[NW],
                                      const uint32_t (&k1)[NW], const uint32_t (&k2)[NW]) {
    uint32_t wd[NW];
    int32_t acc[3];
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].x;
    acc[0] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].y;
    acc[1] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].z;
    acc[2] = tap_sum<NW>(wd, k0, k1, k2);
    sts32(cpx, pack2_clip(acc[0], acc[1], pack2_clip(acc[2], 255 << kPrecisionBits, 0u)));
}
// `replace`: the stand-alone resampler (Image.resize, no over): the un-premultiplied pixel is stored whatever its
// alpha -- Convert.c rgba2rgbA keeps the colours of a pixel whose alpha resampled to 0.
template <int NW>
__device__ __forceinline__ void vcol4(const uint4 (&v)[NW], uint32_t cpx, const uint32_t (&k0)[NW],
                                      const uint32_t (&k1)[NW], const uint32_t (&k2)[NW], bool replace) {
    uint32_t wd[NW];
    int32_t acc[4];
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].w;
    acc[3] = tap_sum<NW>(wd, k0, k1, k2);
    if (acc[3] < (1 << kPrecisionBits) && !replace) return;
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].x;
    acc[0] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].y;
    acc[1] = tap_sum<NW>(wd, k0, k1, k2);
#pragma unroll
    for (int i = 0; i < NW; ++i) wd[i] = v[i].z;
    acc[2] = tap_sum<NW>(wd, k0, k1, k2);
    uint32_t s = pack2_clip(acc[0], acc[1], pack2_clip(acc[2], acc[3], 0u));
    if (replace) s = unpremultiply_px(s);
    else if (acc[3] < (255 << kPrecisionBits)) s = over_unpremul_px(lds32(cpx), s);
    sts32(cpx, s);
}

template <int NW, int NCH>
__device__ __forceinline__ void slab_vpass(const uint32_t *__restrict__ Iw, int CS, uint32_t *__restrict__ ct, int xa, int xb,
                                           int col0, int rw0, int oy0, int tho, int dy, double scale, double support,
                                           const uint32_t *__restrict__ ply, bool replace) {
    const int lane = threadIdx.x & 31;
    uint32_t CSb = 4u * (uint32_t)CS;
    asm volatile("" : "+r"(CSb));
#pragma unroll 1
    for (int r0 = 0; r0 < tho; r0 += 32) {
        const int lrow = min(r0 + lane, tho - 1);
        const int y = oy0 + lrow;
#if B200COMP_ABLATE & 8  // no window start, no coefficient row
        const int wb4 = 0;
        uint32_t k0[NW], k1[NW], k2[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) k0[i] = k1[i] = k2[i] = (uint32_t)y * 0x01010101u;
        (void)scale; (void)support; (void)ply; (void)rw0;
#else
        const int wb4 = ((first_tap(y, scale, support) >> 2) - rw0) * 4;
        uint32_t k0[NW], k1[NW], k2[NW];
        load_coef_row<NW>(ply, y, k0, k1, k2);
#endif
        const int r = dy + lrow;
        const uint32_t crow = smem_u32(ct) + 4u * (uint32_t)((r << 5) + ((col0 & 32) ? kTileH * 32 : 0));
        const uint32_t p_lo = crow + 4u * (uint32_t)((((col0 >> 2) ^ r) & 7) << 2);
        const uint32_t p_hi = crow + 4u * (uint32_t)(((((col0 >> 2) | 1) ^ r) & 7) << 2);
        uint32_t col = smem_u32(Iw) + 4u * (uint32_t)((xa - col0) * CS + wb4);
#pragma unroll 1
        for (int xx = xa - col0; xx < xb - col0; ++xx, col += CSb) {
            uint4 v[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) v[i] = lds128(col + 16u * i);
            const uint32_t cpx = ((xx & 4) ? p_hi : p_lo) + 4u * (uint32_t)(xx & 3);
#if B200COMP_ABLATE & 4  // loads, votes and addresses, no taps / over / store
            uint32_t sink = cpx;
#pragma unroll
            for (int i = 0; i < NW; ++i) sink ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w ^ k0[i] ^ k1[i] ^ k2[i];
            if (__any_sync(0xffffffffu, sink == 0x12345u)) sts32(cpx, sink);
            continue;
#endif
            if (NCH == 3) {
                vcol3<NW>(v, cpx, k0, k1, k2);
            } else {
                uint32_t any = 0u, all = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    any |= v[i].w;
                    all &= v[i].w;
                }
                if (!__any_sync(0xffffffffu, any != 0u)) continue;
                if (__all_sync(0xffffffffu, all == 0xffffffffu)) vcol3<NW>(v, cpx, k0, k1, k2);
                else vcol4<NW>(v, cpx, k0, k1, k2, replace);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
composite_slab_kernel(const Cmd *__restrict__ streams, const int64_t *__restrict__ stream_off,
                      const int64_t *__restrict__ stream_len, const DevCanvas *__restrict__ canvases,
                      const uint8_t *__restrict__ maps, const uint32_t *__restrict__ tables, int slot_words, int iw_words,
                      int n_ring, int *__restrict__ status, uint32_t *__restrict__ dbg) {
    // tile buffers need 1024-byte alignment (swizzle atom): asked of the launch, no slack reserved by hand
    extern __shared__ __align__(1024) uint32_t smem_raw[];
    uint32_t *ctile = smem_raw;
    uint32_t *P = ctile + kTileBufs * kTileWords;  // n_ring slots of slot_words (a multiple of 32 words)
    uint32_t *Iw_all = P + n_ring * slot_words;    // one private intermediate of iw_words per compute warp
    Cmd *ring = reinterpret_cast<Cmd *>(Iw_all + kSlabWarps * iw_words);
    uint32_t *trec = reinterpret_cast<uint32_t *>(ring + kCmdRing * kCmdBlk);  // TILE record of each resident tile
    SlabBars *bars = reinterpret_cast<SlabBars *>(trec + kTileBufs * 16);
    volatile int *n_tiles_total = reinterpret_cast<volatile int *>(bars + 1);  // set by the producer at END

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const Watch watch{status, dbg};
    if (tid == 0) {
        for (int i = 0; i < kPRingMax; ++i) {
            mbar_init(&bars->p_full[i], 1);
            mbar_init(&bars->p_empty[i], kSlabWarps);
        }
        for (int i = 0; i < kTileBufs; ++i) {
            mbar_init(&bars->t_ready[i], 1);
            mbar_init(&bars->t_done[i], kSlabWarps);
            mbar_init(&bars->t_free[i], 1);
        }
        for (int i = 0; i < kCmdRing; ++i) {
            mbar_init(&bars->c_full[i], 1);
            mbar_init(&bars->c_empty[i], kSlabWarps + 1);
        }
        *n_tiles_total = -1;
        mbar_init_fence();
    }
    __syncthreads();
    const Cmd *stream = streams + stream_off[blockIdx.x];

    if (warp == kProducerWarp) {
        const int64_t len = stream_len[blockIdx.x];  // records, END included
        const int nblk = (int)((len + kCmdBlk - 1) / kCmdBlk);
        auto load_block = [&](int b) {
            const int s = b % kCmdRing;
            if (b >= kCmdRing)
                mbar_wait<kProducerSleepNs>(&bars->c_empty[s], ((b / kCmdRing) - 1) & 1, watch, kTagRoleProducer | kTagCmdEmpty, (uint32_t)b);
            if (lane == 0) {
                fence_async_smem();
                mbar_expect_tx(&bars->c_full[s], kCmdBlk * (uint32_t)sizeof(Cmd));
                bulk_load_1d(ring + s * kCmdBlk, stream + (int64_t)b * kCmdBlk, kCmdBlk * (uint32_t)sizeof(Cmd), &bars->c_full[s]);
            }
        };
        if (nblk > 0) load_block(0);
        if (nblk > 1) load_block(1);
        int tseq = 0;
        RingPos rp;
        for (int pos = 0;; ++pos) {
            const int b = pos / kCmdBlk, s = b % kCmdRing, e = pos % kCmdBlk;
            if (e == 0) {
                if (b + 2 < nblk) load_block(b + 2);
                mbar_wait<kProducerSleepNs>(&bars->c_full[s], (b / kCmdRing) & 1, watch, kTagRoleProducer | kTagCmdFull, (uint32_t)pos);
            }
            const Cmd &c = ring[s * kCmdBlk + e];
            const uint32_t kind = uni(c.w[0]);
            if (kind == kCmdEnd) break;
#if B200COMP_L2_PREFETCH
            // The record after this one, if its block is already here: start moving what it will load towards L2.  The
            // ring holds about a third of a step's patch, so the real loads are issued well under a DRAM round trip
            // before they are needed; from L2 they arrive in time.
            if (lane == 0) {
                const int pn = pos + 1, bn = pn / kCmdBlk, sn = bn % kCmdRing;
                if (bn == b || (bn < nblk && mbar_try(&bars->c_full[sn], (bn / kCmdRing) & 1))) {
                    const Cmd &n = ring[sn * kCmdBlk + pn % kCmdBlk];
                    if (n.w[0] == kCmdResample) {
                        const void *map = maps + ((uint64_t)n.w[8] << 7);
                        const int nrq = (int)(n.w[1] >> 24), cw = 4 * (int)(n.w[5] & 0xffffu), rw = (int)(n.w[5] >> 16);
                        for (int q0 = 0; q0 < nrq; q0 += kChunkQuads) tma_prefetch_3d(map, cw, 0, rw + q0);
                    } else if (n.w[0] == kCmdTile && (n.w[6] & kTileBgTma)) {
                        const void *map = reinterpret_cast<const void *>((uint64_t)n.w[8] | ((uint64_t)n.w[9] << 32));
                        tma_prefetch_2d(map, (int)n.w[2], (int)n.w[3]);
                        if ((int)(n.w[4] & 0xffffu) > 32) tma_prefetch_2d(map, (int)n.w[2] + 32, (int)n.w[3]);
                    }
                }
            }
#endif
            if (kind == kCmdTile) {
                const int buf = tseq % kTileBufs;
                if (tseq >= kTileBufs)
                    mbar_wait<kProducerSleepNs>(&bars->t_free[buf], ((tseq / kTileBufs) - 1) & 1, watch, kTagRoleProducer | kTagTileFree, (uint32_t)tseq);
                if (lane < 16) trec[buf * 16 + lane] = c.w[lane];
                __syncwarp();
                if (lane == 0) {
                    if (c.w[6] & kTileBgTma) {
                        const int tw = (int)(c.w[4] & 0xffffu);
                        const void *map = reinterpret_cast<const void *>((uint64_t)c.w[8] | ((uint64_t)c.w[9] << 32));
                        uint32_t *dst = ctile + buf * kTileWords;
                        fence_async_smem();
                        mbar_expect_tx(&bars->t_ready[buf], (tw > 32 ? 2u : 1u) * (uint32_t)(kTileH * 128));
                        tma_load_2d(dst, map, (int)c.w[2], (int)c.w[3], &bars->t_ready[buf]);
                        if (tw > 32) tma_load_2d(dst + kTileH * 32, map, (int)c.w[2] + 32, (int)c.w[3], &bars->t_ready[buf]);
                    } else {
                        mbar_arrive(&bars->t_ready[buf]);
                    }
                }
                ++tseq;
            } else if (kind == kCmdResample || kind == kCmdIdentTma) {
                const bool resample = kind == kCmdResample;
                const uint32_t w1 = uni(c.w[1]), w2 = uni(c.w[2]), w5 = uni(c.w[5]);
                const int dy = (int)((w2 >> 8) & 0xffu), tho = (int)(w2 >> 24);
                const int c_lo = resample ? 0 : dy / kIdentRows;
                const int c_hi = resample ? ((int)(w1 >> 24) + kChunkQuads - 1) / kChunkQuads : (dy + tho - 1) / kIdentRows + 1;
                const void *map = maps + ((uint64_t)uni(c.w[8]) << 7);
                const uint32_t bytes = resample ? uni(c.w[11]) * (uint32_t)(kChunkQuads * 64)
                                                : (uint32_t)(kOverlayBoxW * kIdentRows * 4);
                for (int ci = c_lo; ci < c_hi; ++ci, ring_next(rp, n_ring)) {
                    const int ps = (int)rp.slot;
                    if (rp.seq >= (uint32_t)n_ring) named_bar_sync(1 + ps, (kSlabWarps + 1) * 32);
#if B200COMP_ABLATE & 16  // no patch traffic: the slot is declared full without a copy
                    if (lane == 0) mbar_arrive(&bars->p_full[ps]);
                    (void)map; (void)bytes; (void)w5;
                    continue;
#endif
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect_tx(&bars->p_full[ps], bytes);
                        if (resample)
                            tma_load_3d(P + ps * slot_words, map, 4 * (int)(w5 & 0xffffu), 0, (int)(w5 >> 16) + ci * kChunkQuads, &bars->p_full[ps]);
                        else
                            tma_load_2d(P + ps * slot_words, map, (int)c.w[3], (int)c.w[4] + ci * kIdentRows, &bars->p_full[ps]);
                    }
                }
            }
            if (e == kCmdBlk - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->c_empty[s]);
            }
        }
        if (lane == 0) *n_tiles_total = tseq;
        return;
    }

    if (warp == kStoreWarp) {
        // ---------------- store thread: finished tiles leave through TMA stores, one bulk group per tile ----------------
        if (lane != 0) return;
        for (int tseq = 0;; ++tseq) {
            const int buf = tseq % kTileBufs;
            const uint32_t par = (uint32_t)(tseq / kTileBufs) & 1u;
            bool have = mbar_try(&bars->t_done[buf], par);
            for (uint32_t spins = 0; !have; ++spins) {
                __nanosleep(kStoreSleepNs);
                if (mbar_try(&bars->t_done[buf], par)) { have = true; break; }
                const int total = *n_tiles_total;
                if (total >= 0 && tseq >= total) break;
                if (spins < (1u << 22)) continue;
                if ((*reinterpret_cast<volatile int *>(status) & kStatusWatchdog) == 0)
                    watchdog_fire(watch, kTagRoleStore | kTagTileDone, par, (uint32_t)tseq);
                break;
            }
            if (!have) break;
            const uint32_t *tr = trec + buf * 16;
            if (tr[6] & kTileOutTma) {
                const void *map = reinterpret_cast<const void *>((uint64_t)tr[10] | ((uint64_t)tr[11] << 32));
                const uint32_t *src = ctile + buf * kTileWords;
                fence_async_smem();
                tma_store_2d(map, (int)tr[2], (int)tr[3], src);
                if ((int)(tr[4] & 0xffffu) > 32) tma_store_2d(map, (int)tr[2] + 32, (int)tr[3], src + kTileH * 32);
            }
            bulk_commit();  // tiles the warps stored themselves are an empty group
            bulk_wait_read<0>();  // the store has read the buffer (this thread has nothing else to do meanwhile)
            mbar_arrive(&bars->t_free[buf]);
        }
        bulk_wait_all();
        return;
    }

    // ---------------- compute warps: one slab of kSlabW columns each ----------------
    const int col0 = warp * kSlabW;
    const int jl = lane & 7, q = lane >> 3;
    const int X = col0 + jl;  // this lane's tile column in the element-wise loops and the H pass
    uint32_t *Iw = Iw_all + warp * iw_words;
    int tseq = -1, steps_left = 0;
    RingPos rp;  // chunk ring position (same sequence as the producer's)
    uint32_t t_flags = 0u;
    bool ready_pending = false;  // the wait for the resident tile is deferred to the first access (the H pass does not need it)
    uint32_t *ct = ctile;
    const uint32_t *tr = trec;
    auto tile_wait = [&]() {
        mbar_wait(&bars->t_ready[tseq % kTileBufs], (uint32_t)(tseq / kTileBufs) & 1u, watch, kTagRoleConsumer | kTagTileReady, (uint32_t)tseq);
        ready_pending = false;
    };
    for (int pos = 0;; ++pos) {
        const int b = pos / kCmdBlk, s = b % kCmdRing, e = pos % kCmdBlk;
        if (e == 0) mbar_wait(&bars->c_full[s], (b / kCmdRing) & 1, watch, kTagRoleConsumer | kTagCmdFull, (uint32_t)pos);
        const Cmd &cmd = ring[s * kCmdBlk + e];
        // Record fields are the same for every lane, but the compiler cannot know that of a shared-memory load:
        // broadcasting them from lane 0 marks them warp-uniform (offsets then live on the uniform datapath).
        const uint32_t kind = uni(cmd.w[0]);
        if (kind == kCmdEnd) break;
        if (kind == kCmdTile) {
            ++tseq;
            const int buf = tseq % kTileBufs;
            ct = ctile + buf * kTileWords;
            tr = trec + buf * 16;
            steps_left = (int)uni(cmd.w[1]);
            t_flags = uni(cmd.w[6]);
            ready_pending = true;
            if (!(t_flags & (kTileBgTma | kTileNoBg))) {
                // solid colour, or a background TMA cannot address: every warp fills its own slab
                tile_wait();
                if (t_flags & kTileHasBg) {
                    const DevCanvas &cv = canvases[uni(cmd.w[7])];
                    const uint8_t *bg = cv.bg;
                    const int64_t pitch = cv.bg_pitch;
                    const int tx0 = (int)uni(cmd.w[2]), ty0 = (int)uni(cmd.w[3]);
                    const int tw = (int)(uni(cmd.w[4]) & 0xffffu), th = (int)(uni(cmd.w[4]) >> 16);
                    if (X < tw)
                        for (int yy = q; yy < th; yy += 4)
                            ct[ct_off(yy, X)] = ld_px(bg, (int64_t)(ty0 + yy) * pitch + (int64_t)(tx0 + X) * 4);
                } else {
                    const uint32_t solid = uni(cmd.w[5]);
                    for (int yy = q; yy < kTileH; yy += 4) ct[ct_off(yy, X)] = solid;
                }
                __syncwarp();
            }
        } else {
            const uint32_t w2 = uni(cmd.w[2]);
            const int dx = (int)(w2 & 0xffu), dy = (int)((w2 >> 8) & 0xffu);
            const int two = (int)((w2 >> 16) & 0xffu), tho = (int)(w2 >> 24);
            const int xa = max(col0, dx), xb = min(col0 + kSlabW, dx + two);  // this warp's columns of the step
            if (kind == kCmdResample) {
                const uint32_t w1 = uni(cmd.w[1]);
                const int nwx = (int)(w1 & 0xffu), nwy = (int)((w1 >> 8) & 0xffu);
                const int nch = (int)((w1 >> 16) & 0x7u), NRQ = (int)(w1 >> 24);
                const bool replace = ((w1 >> 23) & 1u) != 0u;  // stand-alone resize: store, do not composite
                if (xa >= xb) {
                    // not on this warp's slab: pass the chunks on
                    for (int q0 = 0; q0 < NRQ; q0 += kChunkQuads, ring_next(rp, n_ring)) {
                        const int ps = (int)rp.slot;
                        mbar_wait(&bars->p_full[ps], rp.phase, watch, kTagRoleConsumer | kTagPatchFull, rp.seq);
                        named_bar_arrive(1 + ps, (kSlabWarps + 1) * 32);
                    }
                } else {
                    const uint32_t w5 = uni(cmd.w[5]);
                    const int ox0 = (int)uni(cmd.w[3]), oy0 = (int)uni(cmd.w[4]);
                    const int cw0 = (int)(w5 & 0xffffu), rw0 = (int)(w5 >> 16);
                    const int n_out_x = (int)uni(cmd.w[6]), n_out_y = (int)uni(cmd.w[7]);
                    const uint32_t *plx = tables + uni(cmd.w[9]), *ply = tables + uni(cmd.w[10]);
                    const int pwc = (int)uni(cmd.w[11]);
                    const double scale_x = __longlong_as_double((long long)((uint64_t)cmd.w[12] | ((uint64_t)cmd.w[13] << 32)));
                    const double scale_y = __longlong_as_double((long long)((uint64_t)cmd.w[14] | ((uint64_t)cmd.w[15] << 32)));
                    // a skipped pass is the 1-tap identity: scale 1, support 1 -> first tap = the sample itself
                    const double support_x = scale_x == 1.0 ? 1.0 : __dmul_rn(3.0, fmax(scale_x, 1.0));
                    const double support_y = scale_y == 1.0 ? 1.0 : __dmul_rn(3.0, fmax(scale_y, 1.0));
                    const int CS = 4 * (NRQ | 1);  // words per intermediate column (odd multiple of 16 bytes)
                    // vertical coefficient rows -> L1 while the horizontal pass runs
                    if (lane < tho) {
                        const uint32_t *row = ply + (int64_t)(oy0 + lane) * coef_row_words(nwy);
                        prefetch_l1(row);
                        prefetch_l1(row + 8);
                    }
                    const bool active = X >= xa && X < xb;
                    const int j = ox0 + (active ? X : xa) - dx;  // lanes off the step read a valid column's table
#define B200_HPASS(NWX, NCH_) \
    slab_hpass<NWX, NCH_>(P, slot_words, n_ring, bars, rp, watch, Iw, CS, NRQ, pwc, cw0, j, active, scale_x, support_x, plx, n_out_x)
#if B200COMP_ABLATE & 1  // the H pass does nothing but hand the chunks on
                    for (int q0 = 0; q0 < NRQ; q0 += kChunkQuads, ring_next(rp, n_ring)) {
                        const int ps = (int)rp.slot;
                        mbar_wait(&bars->p_full[ps], rp.phase, watch, kTagRoleConsumer | kTagPatchFull, rp.seq);
                        named_bar_arrive(1 + ps, (kSlabWarps + 1) * 32);
                    }
                    (void)j; (void)cw0; (void)pwc; (void)plx; (void)n_out_x; (void)support_x;
#else
                    if (nch == 4) {
                        if (nwx == 3) B200_HPASS(3, 4);
                        else if (nwx == 4) B200_HPASS(4, 4);
                        else B200_HPASS(5, 4);
                    } else {
                        if (nwx == 3) B200_HPASS(3, 3);
                        else if (nwx == 4) B200_HPASS(4, 3);
                        else B200_HPASS(5, 3);
                    }
#endif
#undef B200_HPASS
                    __syncwarp();  // the slab's intermediate is complete
                    if (ready_pending) tile_wait();
#define B200_VPASS(NWY, NCH_) \
    slab_vpass<NWY, NCH_>(Iw, CS, ct, xa, xb, col0, rw0, oy0, tho, dy, scale_y, support_y, ply, replace)
#if B200COMP_ABLATE & 2  // no V pass, no over
                    (void)rw0; (void)n_out_y; (void)support_y; (void)replace;
#else
                    if (nch == 4) {
                        if (nwy == 3) B200_VPASS(3, 4);
                        else if (nwy == 4) B200_VPASS(4, 4);
                        else B200_VPASS(5, 4);
                    } else {
                        if (nwy == 3) B200_VPASS(3, 3);
                        else if (nwy == 4) B200_VPASS(4, 3);
                        else B200_VPASS(5, 3);
                    }
#endif
#undef B200_VPASS
                    __syncwarp();  // the next step's H pass overwrites the intermediate
                }
            } else if (kind == kCmdIdentTma) {
                // identity-size overlay: chunks of kIdentRows tile rows of the raw overlay (zero outside it)
                const int shift = (int)uni(cmd.w[5]);
                if (ready_pending) tile_wait();
                for (int ci = dy / kIdentRows; ci <= (dy + tho - 1) / kIdentRows; ++ci, ring_next(rp, n_ring)) {
                    const int ps = (int)rp.slot;
                    mbar_wait(&bars->p_full[ps], rp.phase, watch, kTagRoleConsumer | kTagPatchFull, rp.seq);
                    const uint32_t *slot = P + ps * slot_words;
                    if (X >= xa && X < xb) {
#pragma unroll
                        for (int k = 0; k < kIdentRows / 4; ++k) {
                            const int row = q + 4 * k, yy = ci * kIdentRows + row;
                            if (yy >= dy && yy < dy + tho) {
                                uint32_t *c = ct + ct_off(yy, X);
                                *c = over_px(*c, slot[row * kOverlayBoxW + X + shift]);
                            }
                        }
                    }
                    named_bar_arrive(1 + ps, (kSlabWarps + 1) * 32);
                }
            } else {  // kCmdIdentLdg: overlay read with plain loads (source not addressable by TMA)
                if (ready_pending) tile_wait();
                const uint8_t *src = reinterpret_cast<const uint8_t *>((uint64_t)cmd.w[12] | ((uint64_t)cmd.w[13] << 32));
                const int64_t spitch = (int64_t)uni(cmd.w[6]);
                const int sx0 = (int)uni(cmd.w[3]), sy0 = (int)uni(cmd.w[4]);
                if (X >= xa && X < xb)
                    for (int yy = q; yy < tho; yy += 4) {
                        const uint32_t sp = ld_px(src, (int64_t)(sy0 + yy) * spitch + (int64_t)(sx0 + X - dx) * 4);
                        uint32_t *c = ct + ct_off(dy + yy, X);
                        *c = over_px(*c, sp);
                    }
                __syncwarp();
            }
            if (--steps_left == 0) {
                // the slab's pixels are final
                if (ready_pending) tile_wait();
                if (t_flags & kTileOutTma) {
                    fence_async_smem();  // generic writes to the tile -> visible to the async proxy (TMA store)
                } else {
                    const DevCanvas &cv = canvases[tr[7]];
                    uint8_t *out = cv.out;
                    const int64_t pitch = cv.out_pitch;
                    const int tx0 = (int)tr[2], ty0 = (int)tr[3];
                    const int tw = (int)(tr[4] & 0xffffu), th = (int)(tr[4] >> 16);
                    if (X < tw)
                        for (int yy = q; yy < th; yy += 4)
                            *reinterpret_cast<uint32_t *>(out + (int64_t)(ty0 + yy) * pitch + (int64_t)(tx0 + X) * 4) = ct[ct_off(yy, X)];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->t_done[tseq % kTileBufs]);
            }
        }
        if (e == kCmdBlk - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->c_empty[s]);
        }
    }
}

}  // namespace b200comp
