// libb200comp.so -- kernels and C ABI of the B200 compositor hot path (include/b200comp.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see build.py).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/b200comp.h"
#include "coeffs.h"
#include "kernels.cuh"

namespace b200comp {

// =====================================================================================
// Kernels
// =====================================================================================

// ---- generic one-axis pass (stand-alone resampler, extreme scales) --------------------
// One thread per output pixel; coefficients from global memory (any ksize).
// axis 0: horizontal (in: in_h x in_w -> out: in_h x out_n), axis 1: vertical.
__global__ void resample_axis_kernel(const uint8_t *__restrict__ in, int64_t in_pitch, int in_w, int in_h,
                                     uint8_t *__restrict__ out, int64_t out_pitch, int out_w, int out_h,
                                     const int32_t *__restrict__ k, const int32_t *__restrict__ bounds, int ks,
                                     int axis, int premul_in, int unpremul_out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_w || y >= out_h) return;
    (void)in_w;
    (void)in_h;
    const int o = axis == 0 ? x : y;
    const int lo = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
    const int32_t *kk = k + (int64_t)o * ks;
    int32_t a0, a1, a2, a3;
    a0 = a1 = a2 = a3 = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) {
        const int64_t off = axis == 0 ? (int64_t)y * in_pitch + (int64_t)(lo + t) * 4
                                      : (int64_t)(lo + t) * in_pitch + (int64_t)x * 4;
        uint32_t p = ld_px(in, off);
        if (premul_in) p = premultiply_px(p);
        mac_px(a0, a1, a2, a3, p, __ldg(kk + t));
    }
    uint32_t r = pack_clip(a0, a1, a2, a3);
    if (unpremul_out) r = unpremultiply_px(r);
    *reinterpret_cast<uint32_t *>(out + (int64_t)y * out_pitch + (int64_t)x * 4) = r;
}

// plain / premultiplying / un-premultiplying copy (passes whose size does not change)
__global__ void convert_copy_kernel(const uint8_t *__restrict__ in, int64_t in_pitch, uint8_t *__restrict__ out,
                                    int64_t out_pitch, int w, int h, int premul, int unpremul) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    uint32_t p = ld_px(in, (int64_t)y * in_pitch + (int64_t)x * 4);
    if (premul) p = premultiply_px(p);
    if (unpremul) p = unpremultiply_px(p);
    *reinterpret_cast<uint32_t *>(out + (int64_t)y * out_pitch + (int64_t)x * 4) = p;
}

// ---- in-place alpha-over of one overlay ------------------------------------------------
__global__ void alpha_over_kernel(uint8_t *__restrict__ canvas, int W, int H, int64_t pitch,
                                  const uint8_t *__restrict__ src, int w, int h, int64_t spitch, int x0, int y0) {
    // grid covers the clipped intersection [cx0,cx1) x [cy0,cy1)
    const int cx0 = max(0, x0), cy0 = max(0, y0);
    const int cx1 = min(W, x0 + w), cy1 = min(H, y0 + h);
    const int cx = cx0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = cy0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= cx1 || cy >= cy1) return;
    uint32_t *d = reinterpret_cast<uint32_t *>(canvas + (int64_t)cy * pitch + (int64_t)cx * 4);
    const uint32_t s = ld_px(src, (int64_t)(cy - y0) * spitch + (int64_t)(cx - x0) * 4);
    *d = over_px(*d, s);
}

// ---- fills -----------------------------------------------------------------------------
__global__ void fill_rows_kernel(uint8_t *__restrict__ dst, int W, int H, int64_t pitch, uint32_t rgba) {
    // generic pitch: one row per blockIdx.y, 32-bit stores
    const int y = blockIdx.y;
    uint32_t *row = reinterpret_cast<uint32_t *>(dst + (int64_t)y * pitch);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) row[x] = rgba;
    (void)H;
}

__global__ void fill_flat_kernel(uint4 *__restrict__ dst, size_t n_vec, uint32_t rgba) {
    // contiguous, 16-byte aligned canvas: 128-bit stores, grid-stride
    const uint4 v = make_uint4(rgba, rgba, rgba, rgba);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = v;
}

// fill_gradient (background_resizing.py:74-97): lut[i] = trunc(f32(1-t)*c1 + f32(t)*c2), t = i/max(1,n-1) in double
__global__ void gradient_lut_kernel(uint32_t *__restrict__ lut, int n, int c1r, int c1g, int c1b, int c2r, int c2g,
                                    int c2b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t = (double)i / (double)max(1, n - 1);
    const float a = (float)(1.0 - t);
    const float b = (float)t;
    const uint32_t r = (uint32_t)(int)__fadd_rn(__fmul_rn(a, (float)c1r), __fmul_rn(b, (float)c2r)) & 0xffu;
    const uint32_t g = (uint32_t)(int)__fadd_rn(__fmul_rn(a, (float)c1g), __fmul_rn(b, (float)c2g)) & 0xffu;
    const uint32_t bl = (uint32_t)(int)__fadd_rn(__fmul_rn(a, (float)c1b), __fmul_rn(b, (float)c2b)) & 0xffu;
    lut[i] = r | (g << 8) | (bl << 16) | 0xff000000u;
}

// work item = (row, chunk of 1024 pixels), one warp per item (grid-stride); 128-bit stores when the canvas rows
// are 16-byte aligned (the LUT always is: it comes from the allocator)
__global__ void __launch_bounds__(256)
gradient_fill_kernel(uint8_t *__restrict__ dst, int W, int H, int64_t pitch, const uint32_t *__restrict__ lut,
                     int horizontal, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const int n_chunk = (W + 1023) >> 10;
    const int64_t n_items = (int64_t)H * n_chunk;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n_items; it += n_warps) {
        const int y = (int)(it / n_chunk), xb = (int)(it - (int64_t)y * n_chunk) << 10;
        const int xe = min(W, xb + 1024);
        uint8_t *row = dst + (int64_t)y * pitch;
        const uint32_t rowv = horizontal ? 0u : __ldg(lut + y);
        int x = xb;
        if (vec_ok) {
            for (x = xb + lane * 4; x + 3 < xe; x += 128) {
                const uint4 v = horizontal ? __ldg(reinterpret_cast<const uint4 *>(lut + x)) : make_uint4(rowv, rowv, rowv, rowv);
                *reinterpret_cast<uint4 *>(row + (int64_t)x * 4) = v;
            }
            x = xb + ((xe - xb) & ~3);  // pixels left over after the last whole group of four
        }
        for (x += lane; x < xe; x += 32)
            *reinterpret_cast<uint32_t *>(row + (int64_t)x * 4) = horizontal ? __ldg(lut + x) : rowv;
    }
}

// ---- masked RGB histogram + median -----------------------------------------------------
// hist layout: [2][3][256] uint64: set 0 = pixels with alpha > 0, set 1 = all pixels; counts[2].
// One launch per set: `all_pixels == 0` fills set 0; `all_pixels == 1` fills set 1 and returns at once unless
// set 0 came out empty (background_resizing.py:15-19 falls back to every pixel only then).
//
// The CTA's histogram has one copy per LANE: word (bin, lane) sits in bank `lane`, so the 32 atomics of a warp
// instruction never share a bank, whatever the pixel values (noise and flat colour alike), and an increment is the
// hardware's ATOMS.POPC.INC.  Measured (tools/mb_atoms.cu): 20.8 lane updates per clock per SM this way against
// 13.5 for random bins of a warp-private histogram; the statistics pass needs 10.3 to read at 60 % of HBM.
// VEC: 16-byte loads, lane = 4 consecutive pixels, kHistUnroll loads in flight per lane; otherwise (base or pitch
// not 16-byte aligned) 4-byte loads, lane = pixel.
#ifndef B200COMP_HIST_BRANCH
#define B200COMP_HIST_BRANCH 0
#endif
constexpr int kHistThreads = 512;
constexpr int kHistUnroll = 4;
constexpr size_t kHistSmem = 3 * 256 * 32 * sizeof(unsigned int);  // 96 KB: two CTAs per SM

__device__ __forceinline__ void hist_add_px(uint32_t smem_lane, uint32_t p, bool valid, unsigned int &cnt) {
    // byte offset of word (bin, lane) = bin * 128 + lane * 4; channel planes 256 bins = 32 KB apart (the immediate
    // of the ATOMS address).  One PRMT (byte -> zero-extended word) and one multiply-add per channel.
    const uint32_t a0 = __byte_perm(p, 0u, 0x4440) * 128u + smem_lane, a1 = __byte_perm(p, 0u, 0x4441) * 128u + smem_lane,
                   a2 = __byte_perm(p, 0u, 0x4442) * 128u + smem_lane;
#if B200COMP_HIST_BRANCH
    if (valid) {
        ++cnt;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a0) : "memory");
        asm volatile("red.shared.add.u32 [%0+32768], 1;" ::"r"(a1) : "memory");
        asm volatile("red.shared.add.u32 [%0+65536], 1;" ::"r"(a2) : "memory");
    }
#else
    // masked-out pixels are scattered: instead of a divergent branch per pixel they add 0
    const uint32_t one = valid ? 1u : 0u;
    cnt += one;
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a0), "r"(one) : "memory");
    asm volatile("red.shared.add.u32 [%0+32768], %1;" ::"r"(a1), "r"(one) : "memory");
    asm volatile("red.shared.add.u32 [%0+65536], %1;" ::"r"(a2), "r"(one) : "memory");
#endif
}

template <bool VEC>
__global__ void __launch_bounds__(kHistThreads, 2)
hist_rgb_kernel(const uint8_t *__restrict__ img, int64_t pitch, int x0, int y0, int x1, int y1,
                unsigned long long *__restrict__ hist, unsigned long long *__restrict__ counts, int all_pixels) {
    if (all_pixels && counts[0] != 0ull) return;
    extern __shared__ __align__(16) unsigned int sh[];  // [3][256][32]
    __shared__ unsigned int scount;
    for (int i = threadIdx.x; i < 3 * 256 * 32; i += blockDim.x) sh[i] = 0u;
    if (threadIdx.x == 0) scount = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kHistThreads / 32;
    const uint32_t smem_lane = (uint32_t)__cvta_generic_to_shared(sh) + (uint32_t)lane * 4u;
    const int rw = x1 - x0, nrows = y1 - y0;
    unsigned int cnt = 0;  // pixels counted by this lane
    // work item = (row, chunk of pixels), one warp per item, grid-stride
    constexpr int kPerLoad = VEC ? 128 : 32;  // pixels one warp-wide load covers
    constexpr int kChunk = kPerLoad * kHistUnroll * (VEC ? 1 : 2);
    const int n_chunk = (rw + kChunk - 1) / kChunk;
    const int64_t n_items = (int64_t)nrows * n_chunk;
    for (int64_t it = (int64_t)blockIdx.x * kWarps + warp; it < n_items; it += (int64_t)gridDim.x * kWarps) {
        const int row = (int)(it / n_chunk), c0 = (int)(it - (int64_t)row * n_chunk) * kChunk;
        const int c1 = min(rw, c0 + kChunk);
        const uint8_t *rp = img + (int64_t)(y0 + row) * pitch + (int64_t)x0 * 4;
        if (VEC) {
            uint4 v[kHistUnroll];
#pragma unroll
            for (int j = 0; j < kHistUnroll; ++j) {
                const int x = c0 + j * 128 + lane * 4;
                v[j] = x < c1 ? __ldg(reinterpret_cast<const uint4 *>(rp + (int64_t)x * 4)) : make_uint4(0u, 0u, 0u, 0u);
            }
            if (c1 - c0 == kChunk) {  // whole chunk (warp-uniform): no per-pixel bounds tests
#pragma unroll
                for (int j = 0; j < kHistUnroll; ++j) {
                    hist_add_px(smem_lane, v[j].x, all_pixels || (v[j].x >> 24) != 0u, cnt);
                    hist_add_px(smem_lane, v[j].y, all_pixels || (v[j].y >> 24) != 0u, cnt);
                    hist_add_px(smem_lane, v[j].z, all_pixels || (v[j].z >> 24) != 0u, cnt);
                    hist_add_px(smem_lane, v[j].w, all_pixels || (v[j].w >> 24) != 0u, cnt);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kHistUnroll; ++j) {
                    const int x = c0 + j * 128 + lane * 4;  // the 16-byte load may reach past c1 inside the row's pitch: masked here
                    hist_add_px(smem_lane, v[j].x, x < c1 && (all_pixels || (v[j].x >> 24) != 0u), cnt);
                    hist_add_px(smem_lane, v[j].y, x + 1 < c1 && (all_pixels || (v[j].y >> 24) != 0u), cnt);
                    hist_add_px(smem_lane, v[j].z, x + 2 < c1 && (all_pixels || (v[j].z >> 24) != 0u), cnt);
                    hist_add_px(smem_lane, v[j].w, x + 3 < c1 && (all_pixels || (v[j].w >> 24) != 0u), cnt);
                }
            }
        } else {
            uint32_t px[2 * kHistUnroll];
#pragma unroll
            for (int j = 0; j < 2 * kHistUnroll; ++j) {
                const int x = c0 + j * 32 + lane;
                px[j] = x < c1 ? ld_px(rp, (int64_t)x * 4) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 2 * kHistUnroll; ++j)
                hist_add_px(smem_lane, px[j], c0 + j * 32 + lane < c1 && (all_pixels || (px[j] >> 24) != 0u), cnt);
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(&scount, cnt);
    __syncthreads();
    unsigned long long *out = hist + (all_pixels ? 768 : 0);
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
        unsigned int v = 0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) v += sh[i * 32 + ((k + i) & 31)];  // rotated: the threads of a warp stay on distinct banks
        if (v) atomicAdd(&out[i], (unsigned long long)v);
    }
    if (threadIdx.x == 0 && scount) atomicAdd(&counts[all_pixels ? 1 : 0], (unsigned long long)scount);
}

// np.median semantics: (v[(N-1)/2] + v[N/2]) / 2 per channel; masked set if it has pixels, else all pixels.
// One warp per channel: 8 bins per lane, warp scan of the lane totals, then the two order statistics.
__global__ void __launch_bounds__(96) median_from_hist_kernel(const unsigned long long *__restrict__ hist,
                                                              const unsigned long long *__restrict__ counts, int32_t *__restrict__ out) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int set = counts[0] > 0 ? 0 : 1;
    const unsigned long long n = counts[set];
    if (n == 0) {
        if (lane == 0) out[c] = 0;
        return;
    }
    const unsigned long long *h = hist + set * 768 + c * 256 + lane * 8;
    unsigned long long cum[8], run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        run += h[k];
        cum[k] = run;
    }
    unsigned long long incl = run;  // inclusive scan of the lane totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned long long before = incl - run;
    const unsigned long long ranks[2] = {(n - 1) / 2, n / 2};
    int found[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        // smallest value v with (pixels <= v) > rank
        int mine = 256;
#pragma unroll
        for (int k = 7; k >= 0; --k)
            if (before + cum[k] > ranks[r]) mine = lane * 8 + k;
        found[r] = __reduce_min_sync(0xffffffffu, mine);
    }
    if (lane == 0) out[c] = (found[0] + found[1]) / 2;
}

}  // namespace b200comp

#include "tile_kernel.cuh"
#include "coeff_kernel.cuh"

namespace b200comp {

// ---- the library's own stream-ordered memory pool -------------------------------------------------------
// Scratch buffers and plans come and go (one plan per chunk in the host-buffer pipeline).  They are taken from a
// pool this library creates per device, with a release threshold that keeps freed blocks cached: a repeated call
// pays a pool lookup, not a device allocation (about a millisecond).  The process's default pool, and whatever
// policy the host application gave it, is left alone; b200comp_trim() hands the cached memory back.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};
static cudaMemPool_t lib_pool(int device) {
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        g_pools[device] = pool;
    }
    return g_pools[device];
}
static cudaError_t lib_malloc_async(void **p, size_t bytes, cudaStream_t st) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = lib_pool(device);
    if (!pool) return cudaMallocAsync(p, bytes, st);  // pools unsupported: the default pool as it is
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

// =====================================================================================
// Host side
// =====================================================================================

static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(B200COMP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    } while (0)

static inline cudaStream_t S(void *stream) { return reinterpret_cast<cudaStream_t>(stream); }

static bool aligned4(const void *p, int64_t pitch) {
    return (reinterpret_cast<uintptr_t>(p) & 3u) == 0 && (pitch & 3) == 0;
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// identity table for a skipped pass: one tap of 1.0 at the same index
static void identity_table(int n, int32_t *k, int32_t *b) {
    for (int i = 0; i < n; ++i) {
        k[i] = 1 << kPrecisionBits;
        b[2 * i] = i;
        b[2 * i + 1] = 1;
    }
}

static int launch_axis(const uint8_t *in, int64_t in_pitch, int in_w, int in_h, uint8_t *out, int64_t out_pitch,
                       int out_w, int out_h, const int32_t *k, const int32_t *b, int ks, int axis, int premul,
                       int unpremul, cudaStream_t st) {
    dim3 blk(32, 8), grd((out_w + 31) / 32, (out_h + 7) / 8);
    if (ks > 0)
        resample_axis_kernel<<<grd, blk, 0, st>>>(in, in_pitch, in_w, in_h, out, out_pitch, out_w, out_h, k, b, ks,
                                                   axis, premul, unpremul);
    else
        convert_copy_kernel<<<grd, blk, 0, st>>>(in, in_pitch, out, out_pitch, out_w, out_h, premul, unpremul);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Two-pass resample with device tables; scratch holds the uint8 intermediate.
static int resample_two_pass(const uint8_t *src, int sw, int sh, int64_t sp, uint8_t *dst, int w, int h, int64_t dp,
                             const int32_t *kx, const int32_t *bx, int ksx, const int32_t *ky, const int32_t *by,
                             int ksy, uint8_t *scratch, int flags, cudaStream_t st) {
    if (w == sw && h == sh) {  // Image.resize identity short-cut: copy, no premultiply round trip
        CUDA_TRY(cudaMemcpy2DAsync(dst, dp, src, sp, (size_t)w * 4, h, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const bool need_h = w != sw, need_v = h != sh;
    int rc;
    if (need_h && need_v) {
        if (!scratch) return fail(B200COMP_EINVAL, "scratch buffer required for a two-pass resample");
        if (flags & B200COMP_VERTICAL_FIRST) {
            rc = launch_axis(src, sp, sw, sh, scratch, (int64_t)sw * 4, sw, h, ky, by, ksy, 1, 1, 0, st);
            if (rc) return rc;
            rc = launch_axis(scratch, (int64_t)sw * 4, sw, h, dst, dp, w, h, kx, bx, ksx, 0, 0, 1, st);
        } else {
            rc = launch_axis(src, sp, sw, sh, scratch, (int64_t)w * 4, w, sh, kx, bx, ksx, 0, 1, 0, st);
            if (rc) return rc;
            rc = launch_axis(scratch, (int64_t)w * 4, w, sh, dst, dp, w, h, ky, by, ksy, 1, 0, 1, st);
        }
        return rc;
    }
    if (need_h) return launch_axis(src, sp, sw, sh, dst, dp, w, h, kx, bx, ksx, 0, 1, 1, st);
    return launch_axis(src, sp, sw, sh, dst, dp, w, h, ky, by, ksy, 1, 1, 1, st);
}

// ---- coefficient tables of one plan / call ------------------------------------------------
// Two device formats share one int32 buffer:
//   legacy : k[out][ks] int32 + bounds[out][2]              (generic one-axis kernels)
//   packed : planes[3*nw][out] (byte planes, SoA)            (fused tile kernel, dp4a)
struct TableRef {
    int64_t k_off = 0;  // legacy: k      | packed: w0
    int64_t b_off = 0;  // legacy: bounds | packed: planes
    int ks = 0;         // legacy: taps   | packed: words per output sample (nw)
};

// words per output sample needed by a table with `ks` taps, bucketed to the kernel's variants
static int packed_words(int ks) {
    const int nw = (3 + ks + 3) >> 2;
    if (nw <= 3) return 3;
    if (nw <= 4) return 4;
    if (nw <= 5) return 5;
    return 0;  // more than 17 taps (downscale beyond 2.66x): generic kernels
}

// Legacy (int32) tap tables of the generic kernels are built with double arithmetic and libm on the host: a few tenths
// of a millisecond for an extreme downscale, more than its two kernels take.  The callers that reach them repeat the
// same geometry (thumbnails of the same cutouts every refine iteration, uploads downscaled to the same side), so the
// last tables are kept: 32 MB at most, least recently used first out, values exactly what build_lanczos_table returns.
class LegacyTableCache {
public:
    // k: out_size * ks taps, bounds: out_size * 2
    void get(int in_size, int out_size, int ks, int32_t *k, int32_t *bounds) {
        const size_t nk = (size_t)out_size * ks, nb = (size_t)out_size * 2;
        const uint64_t key = ((uint64_t)(uint32_t)in_size << 32) | (uint32_t)out_size;
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = map_.find(key);
            if (it != map_.end() && it->second.data.size() == nk + nb) {
                it->second.stamp = ++clock_;
                std::memcpy(k, it->second.data.data(), nk * sizeof(int32_t));
                std::memcpy(bounds, it->second.data.data() + nk, nb * sizeof(int32_t));
                return;
            }
        }
        build_lanczos_table(in_size, out_size, k, bounds);
        Entry e;
        e.data.resize(nk + nb);
        std::memcpy(e.data.data(), k, nk * sizeof(int32_t));
        std::memcpy(e.data.data() + nk, bounds, nb * sizeof(int32_t));
        const size_t bytes = e.data.size() * sizeof(int32_t);
        if (bytes > kMaxBytes / 4) return;  // a single huge table would evict everything else
        std::lock_guard<std::mutex> lock(mu_);
        e.stamp = ++clock_;
        bytes_ += bytes;
        auto ins = map_.emplace(key, std::move(e));
        if (!ins.second) bytes_ -= bytes;  // another thread was faster
        while (bytes_ > kMaxBytes && map_.size() > 1) {
            auto oldest = map_.begin();
            for (auto it = map_.begin(); it != map_.end(); ++it)
                if (it->second.stamp < oldest->second.stamp) oldest = it;
            bytes_ -= oldest->second.data.size() * sizeof(int32_t);
            map_.erase(oldest);
        }
    }
    static LegacyTableCache &instance() {
        static LegacyTableCache *c = new LegacyTableCache();  // never destroyed (used from any thread until exit)
        return *c;
    }

private:
    struct Entry {
        std::vector<int32_t> data;
        uint64_t stamp = 0;
    };
    static constexpr size_t kMaxBytes = (size_t)32 << 20;
    std::mutex mu_;
    std::unordered_map<uint64_t, Entry> map_;
    size_t bytes_ = 0;
    uint64_t clock_ = 0;
};

struct TableSet {
    struct Key {
        int in_size, out_size, kind;  // kind 0 legacy, 1 packed, 2 packed identity (skipped pass)
        bool operator<(const Key &o) const {
            return std::tie(in_size, out_size, kind) < std::tie(o.in_size, o.out_size, o.kind);
        }
    };
    std::map<Key, TableRef> refs;
    std::vector<Key> order;
    std::vector<int32_t> host;
    int64_t total = 0;

    TableRef &want(const Key &key) {
        auto it = refs.find(key);
        if (it != refs.end()) return it->second;
        TableRef r;
        const int n = key.out_size;
        if (key.kind == 0) {
            r.ks = lanczos_ksize(key.in_size, key.out_size);
            r.k_off = total;
            total += (int64_t)n * r.ks;
            r.b_off = total;
            total += (int64_t)n * 2;
        } else {
            r.ks = key.kind == 2 ? 3 : packed_words(lanczos_ksize(key.in_size, key.out_size));
            r.k_off = total;
            r.b_off = total;
            total += (int64_t)coef_row_words(r.ks) * n;
        }
        total = (total + 3) & ~(int64_t)3;  // keep every table 16-byte aligned
        order.push_back(key);
        return refs.emplace(key, r).first->second;
    }
    TableRef &want_legacy(int in_size, int out_size) { return want(Key{in_size, out_size, 0}); }
    TableRef &want_packed(int in_size, int out_size, bool identity) {
        return want(Key{in_size, out_size, identity ? 2 : 1});
    }

    // split Pillow's 22-bit taps into byte planes placed at their position inside 4-sample words
    static void pack(int n_out, int ks, const int32_t *k, const int32_t *bounds, int nw, int32_t *planes) {
        uint32_t *pl = reinterpret_cast<uint32_t *>(planes);
        for (int j = 0; j < n_out; ++j) {
            const int lo = bounds[2 * j], n = bounds[2 * j + 1];
            for (int t = 0; t < n; ++t) {
                const int pos = (lo & 3) + t, word = pos >> 2, sh = 8 * (pos & 3);
                const int32_t kv = k[(size_t)j * ks + t];
                uint32_t *row = pl + (size_t)j * coef_row_words(nw);
                row[0 * nw + word] |= (uint32_t)(kv & 0xff) << sh;
                row[1 * nw + word] |= (uint32_t)((kv >> 8) & 0xff) << sh;
                row[2 * nw + word] |= (uint32_t)((kv >> 16) & 0xff) << sh;  // signed top byte
            }
        }
    }

    // skip_packed: the packed tables are built on the device (build_packed_on_device); only the legacy
    // tables of the generic kernels are made here
    bool has_legacy() const {
        for (const Key &k : order) if (k.kind == 0) return true;
        return false;
    }
    void build(int n_threads, bool skip_packed = false) {
        host.assign((size_t)total + 4, 0);
        if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
        n_threads = std::max(1, std::min<int>(n_threads, (int)order.size()));
        auto work = [&](int tid) {
            std::vector<int32_t> k, b;
            for (size_t i = tid; i < order.size(); i += n_threads) {
                const Key &key = order[i];
                const TableRef &r = refs[key];
                int32_t *h = host.data();
                if (skip_packed && key.kind != 0) continue;
                if (key.kind == 0) {
                    LegacyTableCache::instance().get(key.in_size, key.out_size, r.ks, h + r.k_off, h + r.b_off);
                    continue;
                }
                const int n = key.out_size;
                int ks;
                if (key.kind == 2) {
                    ks = 1;
                    k.assign((size_t)n, 0);
                    b.assign((size_t)2 * n, 0);
                    identity_table(n, k.data(), b.data());
                } else {
                    ks = lanczos_ksize(key.in_size, key.out_size);
                    k.assign((size_t)n * ks, 0);
                    b.assign((size_t)2 * n, 0);
                    build_lanczos_table(key.in_size, key.out_size, k.data(), b.data());
                }
                pack(n, ks, k.data(), b.data(), r.ks, h + r.b_off);
            }
        };
        if (n_threads == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
            for (auto &t : th) t.join();
        }
    }
};

// Packed tables on the device (coeff_kernel.cuh): one launch computes every output sample of every table;
// samples whose fixed-point value sits within the guard band of a rounding boundary come back in a list
// and are recomputed here with libm, so the result is bit-identical to the host builder.
// Returns 0, or 1 when the fix-up list overflowed (caller falls back to the host builder).
static int build_packed_on_device(const TableSet &ts, int32_t *d_tables, cudaStream_t st, int64_t *n_fixed, std::string *err) {
    std::vector<CoefJob> jobs;
    int max_out = 1;
    int64_t samples = 0;
    for (const TableSet::Key &key : ts.order) {
        if (key.kind == 0) continue;
        const TableRef &r = ts.refs.at(key);
        CoefJob j;
        j.planes_off = r.b_off;
        j.in_size = key.in_size;
        j.out_size = key.out_size;
        j.nw = r.ks;
        j.identity = key.kind == 2 ? 1 : 0;
        jobs.push_back(j);
        max_out = std::max(max_out, key.out_size);
        samples += key.out_size;
    }
    if (n_fixed) *n_fixed = 0;
    if (jobs.empty()) return 0;
    const int fix_cap = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(4096, samples / 64));
    CoefJob *d_jobs = nullptr;
    CoefFix *d_fix = nullptr;
    int *d_count = nullptr;
    auto fail_cuda = [&](cudaError_t e) {
        if (err) *err = std::string("device coefficient tables: ") + cudaGetErrorString(e);
        if (d_jobs) cudaFreeAsync(d_jobs, st);
        if (d_fix) cudaFreeAsync(d_fix, st);
        if (d_count) cudaFreeAsync(d_count, st);
        return -1;
    };
    cudaError_t e;
    if ((e = lib_malloc_async((void **)&d_jobs, jobs.size() * sizeof(CoefJob), st)) != cudaSuccess) return fail_cuda(e);
    if ((e = lib_malloc_async((void **)&d_fix, (size_t)fix_cap * sizeof(CoefFix), st)) != cudaSuccess) return fail_cuda(e);
    if ((e = lib_malloc_async((void **)&d_count, sizeof(int), st)) != cudaSuccess) return fail_cuda(e);
    if ((e = cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(CoefJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail_cuda(e);
    if ((e = cudaMemsetAsync(d_count, 0, sizeof(int), st)) != cudaSuccess) return fail_cuda(e);
    const unsigned gx = (unsigned)std::min(16, (max_out + 127) / 128);
    for (size_t j0 = 0; j0 < jobs.size(); j0 += 65535) {
        const unsigned ny = (unsigned)std::min<size_t>(65535, jobs.size() - j0);
        build_packed_tables_kernel<<<dim3(gx, ny), 128, 0, st>>>(d_jobs + j0, (int)j0, reinterpret_cast<uint32_t *>(d_tables), d_fix,
                                                                 d_count, fix_cap);
    }
    int count = 0;
    if ((e = cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda(e);
    int rc = 0;
    if (count > fix_cap) {
        rc = 1;  // too many borderline samples for the list: let the host build everything
    } else if (count > 0) {
        std::vector<CoefFix> fix((size_t)count);
        if ((e = cudaMemcpyAsync(fix.data(), d_fix, (size_t)count * sizeof(CoefFix), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail_cuda(e);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail_cuda(e);
        std::vector<WordPatch> patches;
        std::vector<int32_t> row;
        for (const CoefFix &f : fix) {
            const CoefJob &job = jobs[(size_t)f.job];
            row.assign((size_t)lanczos_ksize(job.in_size, job.out_size) + 1, 0);
            int lo = 0, n = 0;
            build_lanczos_row(job.in_size, job.out_size, f.j, row.data(), &lo, &n);
            uint32_t pl[3][5] = {{0}};
            for (int t = 0; t < n; ++t) {
                const int pos = (lo & 3) + t, word = pos >> 2, sh = 8 * (pos & 3);
                if (word >= 5) break;
                pl[0][word] |= (uint32_t)(row[(size_t)t] & 0xff) << sh;
                pl[1][word] |= (uint32_t)((row[(size_t)t] >> 8) & 0xff) << sh;
                pl[2][word] |= (uint32_t)((row[(size_t)t] >> 16) & 0xff) << sh;
            }
            for (int p = 0; p < 3; ++p)
                for (int i = 0; i < job.nw; ++i) {
                    WordPatch wp;
                    wp.off = job.planes_off + (int64_t)f.j * coef_row_words(job.nw) + p * job.nw + i;
                    wp.value = pl[p][i];
                    wp.pad_ = 0;
                    patches.push_back(wp);
                }
        }
        WordPatch *d_patches = nullptr;
        if ((e = lib_malloc_async((void **)&d_patches, patches.size() * sizeof(WordPatch), st)) != cudaSuccess) return fail_cuda(e);
        e = cudaMemcpyAsync(d_patches, patches.data(), patches.size() * sizeof(WordPatch), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            patch_words_kernel<<<(unsigned)((patches.size() + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint32_t *>(d_tables),
                                                                                        d_patches, (int)patches.size());
            e = cudaStreamSynchronize(st);  // `patches` is a local
        }
        cudaFreeAsync(d_patches, st);
        if (e != cudaSuccess) return fail_cuda(e);
        if (n_fixed) *n_fixed = count;
    }
    cudaFreeAsync(d_jobs, st);
    cudaFreeAsync(d_fix, st);
    cudaFreeAsync(d_count, st);
    return rc;
}

// Upper bound of the 4-sample words one tile needs along an axis (see tile_kernel.cuh: cw1 - cw0).
static int words_bound(int in_size, int out_size, int n_out, int nw) {
    const double scale = (double)in_size / out_size;
    const double span = (std::min(n_out, out_size) - 1) * scale + 1.0;
    return (int)std::floor(span / 4.0) + 2 + nw;
}

}  // namespace b200comp

using namespace b200comp;

// =====================================================================================
// Plan
// =====================================================================================
static constexpr int kMaxWaves = 16;  // sub-ranges of canvases a run is cut into (see b200comp_plan_run_canvases)

struct b200comp_plan {
    int device = 0;
    cudaStream_t create_stream = nullptr;
    int n_canvases = 0;
    int64_t n_tiles = 0;
    int max_tiles = 1;
    int32_t *d_tables = nullptr;
    DevPlacementT *d_placements = nullptr;
    DevCanvas *d_canvases = nullptr;
    int *d_status = nullptr;
    int slot_words = 0, iw_words = 0;  // ring slot of the widest patch class / private intermediate of one warp
    int n_ring = kPRingMin;            // slots of the patch chunk ring (as many as keep kCtasPerSm CTAs per SM resident)
    size_t smem_bytes = 0;
    int64_t info[B200COMP_INFO_COUNT] = {0};
    // placements resampled by the generic kernels before the tile kernel (extreme scales, vertical-first)
    struct Pre {
        const uint8_t *src;
        int64_t sp;
        int sw, sh, w, h, flags;
        TableRef tx, ty;
        uint8_t *dst;      // w*h*4 temp
        uint8_t *scratch;  // intermediate
    };
    std::vector<Pre> pre;
    PrepDesc *d_prep = nullptr;  // distinct cutouts the tile kernel resamples (prepared every run)
    uint32_t *d_flags = nullptr; // alpha summaries of the prepared cutouts
    size_t flag_bytes = 0;
    int n_prep = 0;
    int prep_blocks_x = 1;
    // command streams of the persistent tile kernel (rebuilt by the binning kernels at every run)
    int G = 1;                       // persistent CTAs = streams
    int64_t stream_capacity = 0;     // records (exact upper bound from the host-side box/tile count)
    Cmd *d_streams = nullptr;
    int32_t *d_bin = nullptr;        // per wave [G][K_wave] per-tile slot counts, scanned in place
    int64_t *d_stream_off = nullptr; // [kMaxWaves][G + 2]: offsets, then the wave's record cursor
    int64_t *d_stream_len = nullptr; // [kMaxWaves][G] records per stream, END included
    uint32_t *d_dbg = nullptr;       // 16 words written by the tile kernel's watchdog
    uint8_t *d_maps = nullptr;       // every CUtensorMap of the plan (placements, overlays, canvases)
    uint32_t *d_masks = nullptr;     // [tiles][mask_chunks][2] keep / opaque masks from the count kernel
    int4 *d_boxes = nullptr;         // destination boxes (x, y, w, h) of the placements: the binning hit test
    int mask_chunks = 1;             // ceil(max placements per canvas / 32)
    std::vector<int64_t> tiles_before;  // prefix sum of tiles per canvas (n_canvases + 1)
    std::vector<int64_t> records_before;  // prefix sum of the record upper bounds per canvas (n_canvases + 1)
    // waves of a run (b200comp_plan_run_canvases): binning of wave w+1 runs on `s_bin` under the tile kernel of
    // wave w; tile kernels alternate between the caller's stream and `s_tile` so one fills the other's tail
    cudaStream_t s_bin = nullptr, s_tile = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_tile = nullptr, ev_bin[kMaxWaves] = {};
    int last_waves = 0;  // waves of the most recent run
    // b200comp_plan_profile: events around the phases of every run (4 per run: before prepare, before
    // binning, before the tile kernel, after it); a prepare without a run contributes nothing
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;
    cudaEvent_t prof_prepare[2] = {nullptr, nullptr};
    bool prof_prepare_valid = false;
    int prof_runs = 0;
    std::vector<void *> owned;  // device allocations freed with the plan
};

// number of waves a run over `count` canvases / `n_tiles` tiles is cut into (b200comp_plan_run_canvases)
static int wave_count(const b200comp_plan *plan, int64_t n_tiles, int count) {
    const char *e = std::getenv("B200COMP_WAVES");  // read at every run: the parity tests compare wave counts
    const int forced = e && e[0] ? std::max(1, std::min(kMaxWaves, std::atoi(e))) : 0;
    if (plan->profile) return 1;  // the phase events of b200comp_plan_profile want the phases one after another
    int n = forced ? forced : (int)std::min<int64_t>(kMaxWaves, n_tiles / (32 * 2048));  // >= ~32 4K canvases per wave
    return std::max(1, std::min(n, count));
}

static const size_t kMaxSmemBytes = 200 * 1024;    // opt-in dynamic shared memory limit we request
// Two CTAs per SM: 228 KB of shared memory less the 1 KB the system reserves per CTA.  Placements needing more (with the
// shallowest ring) go through the generic kernels.
static const size_t kFusedSmemCap = 113 * 1024;
// dynamic shared memory of the tile kernel: resident tiles + patch chunk ring + one private intermediate per compute
// warp + command blocks + TILE records + mbarriers
static size_t tile_smem_bytes(int64_t slot_words, int64_t iw_words, int n_ring = kPRingMin) {
    return ((size_t)kTileBufs * kTileWords + (size_t)n_ring * slot_words + (size_t)kSlabWarps * iw_words) * 4 +
           (size_t)kCmdRing * kCmdBlk * sizeof(Cmd) + (size_t)kTileBufs * 16 * sizeof(uint32_t) + sizeof(SlabBars) + 16;
}
// ring slot of a placement whose patch rows are `pwc` word columns wide: kChunkQuads quads x 4 planes x 4 * pwc words
static int slot_words_of(int pwc) { return kChunkQuads * 16 * pwc; }
static_assert(sizeof(b200comp_placement) == 48 && sizeof(b200comp_canvas) == 56, "public struct layout");

#pragma GCC visibility push(default)
extern "C" {

int b200comp_abi_version(void) { return B200COMP_ABI_VERSION; }
// internal helpers shared with host_api.cu (not part of the public header): the library pool
int b200comp_pool_alloc_(void **p, size_t bytes, void *stream) { return (int)lib_malloc_async(p, bytes, S(stream)); }
void b200comp_pool_trim_(void) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (g_pools[device]) cudaMemPoolTrimTo(g_pools[device], 0);
}
// internal helper shared with host_api.cu (not part of the public header)
int b200comp_set_error_(int code, const char *msg) { return fail(code, msg ? msg : ""); }
const char *b200comp_last_error(void) { return g_err.c_str(); }

int b200comp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200comp_ksize(int in_size, int out_size) {
    if (in_size < 1 || out_size < 1) return fail(B200COMP_EINVAL, "sizes must be positive");
    return lanczos_ksize(in_size, out_size);
}

int b200comp_build_coeffs(int in_size, int out_size, int32_t *k_host, int32_t *bounds_host, int *ksize) {
    if (in_size < 1 || out_size < 1 || !k_host || !bounds_host)
        return fail(B200COMP_EINVAL, "b200comp_build_coeffs: bad argument");
    const int ks = build_lanczos_table(in_size, out_size, k_host, bounds_host);
    if (ksize) *ksize = ks;
    return 0;
}

int b200comp_resample_rgba(const uint8_t *src, int sw, int sh, size_t src_pitch, uint8_t *dst, int w, int h,
                           size_t dst_pitch, const int32_t *kx, const int32_t *bx, int ksx, const int32_t *ky,
                           const int32_t *by, int ksy, uint8_t *scratch, int flags, void *stream) {
    if (!src || !dst || sw < 1 || sh < 1 || w < 1 || h < 1) return fail(B200COMP_EINVAL, "resample: bad size/pointer");
    if (!aligned4(src, (int64_t)src_pitch) || !aligned4(dst, (int64_t)dst_pitch))
        return fail(B200COMP_EINVAL, "resample: buffers must be 4-byte aligned with pitch % 4 == 0");
    if ((w != sw && (!kx || !bx || ksx < 1)) || (h != sh && (!ky || !by || ksy < 1)))
        return fail(B200COMP_EINVAL, "resample: missing coefficient table");
    return resample_two_pass(src, sw, sh, (int64_t)src_pitch, dst, w, h, (int64_t)dst_pitch, kx, bx, ksx, ky, by, ksy,
                             scratch, flags, S(stream));
}

int b200comp_resize_rgba_lanczos(const uint8_t *src, int sw, int sh, size_t src_pitch, uint8_t *dst, int w, int h,
                                 size_t dst_pitch, int flags, void *stream) {
    if (!src || !dst || sw < 1 || sh < 1 || w < 1 || h < 1) return fail(B200COMP_EINVAL, "resize: bad size/pointer");
    if (!aligned4(src, (int64_t)src_pitch) || !aligned4(dst, (int64_t)dst_pitch))
        return fail(B200COMP_EINVAL, "resize: buffers must be 4-byte aligned with pitch % 4 == 0");
    cudaStream_t st = S(stream);
    if (w == sw && h == sh)
        return resample_two_pass(src, sw, sh, (int64_t)src_pitch, dst, w, h, (int64_t)dst_pitch, nullptr, nullptr, 0,
                                 nullptr, nullptr, 0, nullptr, flags, st);
    // The fused tile kernel in "replace" mode onto a canvas that is the destination: prepared planar source, TMA patch
    // chunks, dp4a passes, device-built coefficient tables (scales down to 2.66x; buffers of any 4-byte alignment).
    // Only what it cannot take -- more taps, Pillow 12's vertical-first order -- goes through the generic kernels below.
    {
        const bool need_h = w != sw, need_v = h != sh;
        const bool vertical_first = (flags & B200COMP_VERTICAL_FIRST) && need_h && need_v;
        const int nwx = need_h ? packed_words(lanczos_ksize(sw, w)) : 3, nwy = need_v ? packed_words(lanczos_ksize(sh, h)) : 3;
        if (!vertical_first && nwx > 0 && nwy > 0 && src_pitch <= (size_t)INT32_MAX && sw < 262144 && sh < 262144) {
            b200comp_canvas cv;
            std::memset(&cv, 0, sizeof cv);
            cv.out = dst;
            cv.out_pitch = (int64_t)dst_pitch;
            cv.W = w;
            cv.H = h;
            cv.n_placements = 1;
            b200comp_placement pl;
            std::memset(&pl, 0, sizeof pl);
            pl.src = src;
            pl.src_pitch = (int64_t)src_pitch;
            pl.sw = sw; pl.sh = sh; pl.w = w; pl.h = h;
            pl.flags = B200COMP_REPLACE;
            b200comp_plan *plan = nullptr;
            const int rc = b200comp_plan_create(&cv, 1, &pl, 1, 1, stream, &plan);
            if (rc == 0) {
                const int rr = b200comp_plan_run(plan, stream);
                b200comp_plan_destroy(plan);  // frees are ordered on `stream`, behind the run
                return rr;
            }
            if (rc != B200COMP_EINVAL) return rc;  // EINVAL: the placement needs more shared memory than the fused kernel has
        }
    }
    TableSet ts;
    TableRef tx, ty;
    if (w != sw) tx = ts.want_legacy(sw, w);
    if (h != sh) ty = ts.want_legacy(sh, h);
    ts.build((int64_t)w + h > 8192 ? 2 : 1);  // two threads only for very long tables (spawning one costs as much as a cached or mid-sized table)
    const size_t tbytes = ts.host.size() * sizeof(int32_t);
    const size_t sbytes = (size_t)std::max((int64_t)sh * w, (int64_t)h * sw) * 4;
    uint8_t *d_mem = nullptr;
    CUDA_TRY(lib_malloc_async((void **)&d_mem, tbytes + sbytes + 16, st));
    int32_t *d_t = reinterpret_cast<int32_t *>(d_mem);
    uint8_t *d_s = d_mem + ((tbytes + 15) & ~(size_t)15);
    cudaError_t e = cudaMemcpyAsync(d_t, ts.host.data(), tbytes, cudaMemcpyHostToDevice, st);
    int rc = 0;
    if (e == cudaSuccess) {
        // the pageable-source copy above is staged before it returns, so ts may go out of scope
        rc = resample_two_pass(src, sw, sh, (int64_t)src_pitch, dst, w, h, (int64_t)dst_pitch, d_t + tx.k_off,
                               d_t + tx.b_off, tx.ks, d_t + ty.k_off, d_t + ty.b_off, ty.ks, d_s, flags, st);
    }
    cudaFreeAsync(d_mem, st);
    if (e != cudaSuccess) return fail(B200COMP_ECUDA, std::string("table upload: ") + cudaGetErrorString(e));
    return rc;
}

int b200comp_alpha_over(uint8_t *canvas, int W, int H, size_t pitch, const uint8_t *src, int w, int h,
                        size_t src_pitch, int x, int y, void *stream) {
    if (!canvas || !src || W < 1 || H < 1 || w < 1 || h < 1) return fail(B200COMP_EINVAL, "alpha_over: bad argument");
    if (!aligned4(canvas, (int64_t)pitch) || !aligned4(src, (int64_t)src_pitch))
        return fail(B200COMP_EINVAL, "alpha_over: buffers must be 4-byte aligned with pitch % 4 == 0");
    const int64_t cx0 = std::max<int64_t>(0, x), cy0 = std::max<int64_t>(0, y);
    const int64_t cx1 = std::min<int64_t>(W, (int64_t)x + w), cy1 = std::min<int64_t>(H, (int64_t)y + h);
    if (cx0 >= cx1 || cy0 >= cy1) return 0;  // fully outside: nothing changes
    dim3 blk(32, 8), grd((unsigned)((cx1 - cx0 + 31) / 32), (unsigned)((cy1 - cy0 + 7) / 8));
    alpha_over_kernel<<<grd, blk, 0, S(stream)>>>(canvas, W, H, (int64_t)pitch, src, w, h, (int64_t)src_pitch, x, y);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200comp_fill_rgba(uint8_t *dst, int W, int H, size_t pitch, uint32_t rgba, void *stream) {
    if (!dst || W < 1 || H < 1) return fail(B200COMP_EINVAL, "fill: bad argument");
    if (!aligned4(dst, (int64_t)pitch)) return fail(B200COMP_EINVAL, "fill: buffer must be 4-byte aligned");
    cudaStream_t st = S(stream);
    const size_t bytes = (size_t)W * 4 * H;
    if (pitch == (size_t)W * 4 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && bytes >= 16) {
        const size_t n_vec = bytes / 16;
        const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
        fill_flat_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4 *>(dst), n_vec, rgba);
        const size_t tail_px = (bytes - n_vec * 16) / 4;
        if (tail_px) fill_rows_kernel<<<dim3(1, 1), 32, 0, st>>>(dst + n_vec * 16, (int)tail_px, 1, 0, rgba);
    } else {
        dim3 grd((unsigned)std::min(16, (W + 255) / 256), (unsigned)H);
        fill_rows_kernel<<<grd, 256, 0, st>>>(dst, W, H, (int64_t)pitch, rgba);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200comp_fill_gradient(uint8_t *dst, int W, int H, size_t pitch, int horizontal, const int32_t c1[3],
                           const int32_t c2[3], void *stream) {
    if (!dst || W < 1 || H < 1 || !c1 || !c2) return fail(B200COMP_EINVAL, "fill_gradient: bad argument");
    if (!aligned4(dst, (int64_t)pitch)) return fail(B200COMP_EINVAL, "fill_gradient: buffer must be 4-byte aligned");
    cudaStream_t st = S(stream);
    const int n = horizontal ? W : H;
    uint32_t *lut = nullptr;
    CUDA_TRY(lib_malloc_async((void **)&lut, (size_t)n * 4, st));
    gradient_lut_kernel<<<(n + 255) / 256, 256, 0, st>>>(lut, n, c1[0], c1[1], c1[2], c2[0], c2[1], c2[2]);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)pitch) & 15u) == 0;
    const int64_t items = (int64_t)H * ((W + 1023) / 1024);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((items + 7) / 8, 148 * 8));
    gradient_fill_kernel<<<blocks, 256, 0, st>>>(dst, W, H, (int64_t)pitch, lut, horizontal ? 1 : 0, vec_ok);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(lut, st);
    if (e != cudaSuccess) return fail(B200COMP_ECUDA, cudaGetErrorString(e));
    return 0;
}

int b200comp_masked_median_rgb(const uint8_t *img, int W, int H, size_t pitch, int x0, int y0, int x1, int y1,
                               int32_t out_rgb[3], void *stream) {
    if (!img || !out_rgb || W < 1 || H < 1) return fail(B200COMP_EINVAL, "median: bad argument");
    if (x0 < 0 || y0 < 0 || x1 > W || y1 > H || x0 >= x1 || y0 >= y1)
        return fail(B200COMP_EINVAL, "median: rectangle outside the image or empty");
    if (!aligned4(img, (int64_t)pitch)) return fail(B200COMP_EINVAL, "median: buffer must be 4-byte aligned");
    cudaStream_t st = S(stream);
    unsigned long long *d = nullptr;  // [2*3*256] hist + [2] counts + 3 int32 result
    const size_t bytes = (2 * 3 * 256 + 2) * sizeof(unsigned long long) + 4 * sizeof(int32_t);
    CUDA_TRY(lib_malloc_async((void **)&d, bytes, st));
    cudaError_t e = cudaMemsetAsync(d, 0, bytes, st);
    int32_t *d_out = reinterpret_cast<int32_t *>(d + 2 * 3 * 256 + 2);
    if (e == cudaSuccess) {
        static const bool smem_set = [] {
            cudaFuncSetAttribute(hist_rgb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistSmem);
            cudaFuncSetAttribute(hist_rgb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistSmem);
            return true;
        }();
        (void)smem_set;
        // 16-byte loads when every row start is 16-byte aligned
        const bool vec = ((reinterpret_cast<uintptr_t>(img) + (uintptr_t)x0 * 4) & 15u) == 0 && (pitch & 15u) == 0;
        const int chunk = vec ? 128 * kHistUnroll : 64 * kHistUnroll;
        const int64_t items = (int64_t)(y1 - y0) * ((x1 - x0 + chunk - 1) / chunk);  // warp = (row, chunk)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int warps = kHistThreads / 32;
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((items + warps - 1) / warps, (int64_t)sms * 2));
        // masked set first; the all-pixel launch returns immediately unless no pixel had alpha > 0
        for (int all = 0; all < 2; ++all) {
            if (vec)
                hist_rgb_kernel<true><<<blocks, kHistThreads, kHistSmem, st>>>(img, (int64_t)pitch, x0, y0, x1, y1, d, d + 2 * 3 * 256, all);
            else
                hist_rgb_kernel<false><<<blocks, kHistThreads, kHistSmem, st>>>(img, (int64_t)pitch, x0, y0, x1, y1, d, d + 2 * 3 * 256, all);
        }
        median_from_hist_kernel<<<1, 96, 0, st>>>(d, d + 2 * 3 * 256, d_out);
        e = cudaGetLastError();
    }
    int32_t h_out[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
    if (e != cudaSuccess) return fail(B200COMP_ECUDA, std::string("median: ") + cudaGetErrorString(e));
    out_rgb[0] = h_out[0];
    out_rgb[1] = h_out[1];
    out_rgb[2] = h_out[2];
    return 0;
}

// -------------------------------------------------------------------------------- plan
int b200comp_plan_destroy(b200comp_plan *plan) {
    if (!plan) return 0;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != plan->device) cudaSetDevice(plan->device);
    for (void *p : plan->owned) cudaFreeAsync(p, plan->create_stream);
    for (cudaEvent_t e : plan->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : plan->prof_prepare) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : plan->ev_bin) if (e) cudaEventDestroy(e);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_tile) cudaEventDestroy(plan->ev_tile);
    if (plan->s_bin) cudaStreamDestroy(plan->s_bin);  // pending work (already joined into the caller's stream) still completes
    if (plan->s_tile) cudaStreamDestroy(plan->s_tile);
    if (cur != plan->device) cudaSetDevice(cur);
    delete plan;
    return 0;
}

int b200comp_plan_create(const b200comp_canvas *canvases, int n_canvases, const b200comp_placement *placements,
                         int n_placements, int n_host_threads, void *stream, b200comp_plan **out_plan) {
    if (!out_plan) return fail(B200COMP_EINVAL, "plan_create: null plan pointer");
    *out_plan = nullptr;
    if (n_canvases < 1 || !canvases || n_placements < 0 || (n_placements > 0 && !placements))
        return fail(B200COMP_EINVAL, "plan_create: empty batch or null arrays");
    cudaStream_t st = S(stream);
    b200comp_plan *plan = new b200comp_plan();
    struct Guard {
        b200comp_plan *p;
        ~Guard() { if (p) b200comp_plan_destroy(p); }
    } guard{plan};
    CUDA_TRY(cudaGetDevice(&plan->device));
    plan->create_stream = st;
    plan->n_canvases = n_canvases;

    // B200COMP_TRACE_PLAN=1: host-side timeline of plan creation on stderr
    static const bool trace_plan = std::getenv("B200COMP_TRACE_PLAN") != nullptr;
    const auto t_trace0 = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (!trace_plan) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_trace0).count();
        std::fprintf(stderr, "[b200comp plan] %8.3f ms  %s\n", ms, what);
    };
    // ---- validate, classify placements, collect tables ----
    TableSet ts;
    std::vector<DevPlacementT> hp((size_t)std::max(1, n_placements));
    std::vector<std::pair<TableRef, TableRef>> tref((size_t)std::max(1, n_placements));
    std::vector<int> pre_index((size_t)std::max(1, n_placements), -1);
    int max_slot = kOverlayBoxW * kIdentRows, max_iw = 16;  // the chunk ring also stages identity overlays
    int64_t n_fused = 0, n_ident = 0;
    for (int i = 0; i < n_placements; ++i) {
        const b200comp_placement &p = placements[i];
        if (!p.src || p.sw < 1 || p.sh < 1 || p.w < 1 || p.h < 1)
            return fail(B200COMP_EINVAL, "plan_create: placement " + std::to_string(i) + " has a null source or empty size");
        if (!aligned4(p.src, p.src_pitch) || p.src_pitch < (int64_t)p.sw * 4 || p.src_pitch > INT32_MAX)
            return fail(B200COMP_EINVAL, "plan_create: placement " + std::to_string(i) + " source misaligned or bad pitch");
        // the device hit tests add x + w and y + h in int32
        if (p.x < -(1 << 30) || p.x > (1 << 30) || p.y < -(1 << 30) || p.y > (1 << 30) || p.w > (1 << 30) || p.h > (1 << 30))
            return fail(B200COMP_EINVAL, "plan_create: placement " + std::to_string(i) + " box out of range (|x|, |y|, w, h <= 2^30)");
        DevPlacementT &d = hp[i];
        std::memset(&d, 0, sizeof d);
        d.src = p.src;
        d.src_pitch = (int32_t)p.src_pitch;
        d.sw = p.sw; d.sh = p.sh;
        d.x = p.x; d.y = p.y; d.w = p.w; d.h = p.h;
        if (p.w == p.sw && p.h == p.sh) {
            if (p.flags & B200COMP_REPLACE)
                return fail(B200COMP_EINVAL, "plan_create: B200COMP_REPLACE needs a placement the fused kernel resamples");
            d.mode = 0;
            ++n_ident;
            continue;
        }
        const bool need_h = p.w != p.sw, need_v = p.h != p.sh;
        const int nwx = need_h ? packed_words(lanczos_ksize(p.sw, p.w)) : 3;
        const int nwy = need_v ? packed_words(lanczos_ksize(p.sh, p.h)) : 3;
        const bool vertical_first = (p.flags & B200COMP_VERTICAL_FIRST) && need_h && need_v;
        bool fused = !vertical_first && nwx > 0 && nwy > 0;
        int pwc = 0, slot = 0, iw = 0;
        if (fused) {
            // patch width class: word columns one tile step can need, rounded up to a multiple of 4 (one tensor map
            // per cutout and class); row quads one tile step can need -> the warps' private intermediates
            pwc = (words_bound(p.sw, p.w, kTileW, nwx) + 3) & ~3;
            const int nrq = words_bound(p.sh, p.h, kTileH, nwy);
            slot = slot_words_of(pwc);
            iw = 4 * kSlabW * (nrq | 1);
            fused = tile_smem_bytes(std::max(slot, kOverlayBoxW * kIdentRows), iw) <= kFusedSmemCap && 4 * pwc <= 256 && nrq <= 255 &&
                    p.sw < 262144 && p.sh < 262144;
        }
        if (fused) {
            d.mode = 1;
            tref[i].first = ts.want_packed(p.sw, p.w, !need_h);
            tref[i].second = ts.want_packed(p.sh, p.h, !need_v);
            d.nwx = tref[i].first.ks;
            d.nwy = tref[i].second.ks;
            d.pwc = pwc;
            d.replace = (p.flags & B200COMP_REPLACE) ? 1 : 0;
            // the kernel recomputes each window start from these doubles exactly as the table builder does;
            // a skipped pass is the 1-tap identity: scale 1, support 1 -> first tap = the sample itself
            d.scale_x = need_h ? (double)p.sw / p.w : 1.0;
            d.support_x = need_h ? 3.0 * std::max(d.scale_x, 1.0) : 1.0;
            d.scale_y = need_v ? (double)p.sh / p.h : 1.0;
            d.support_y = need_v ? 3.0 * std::max(d.scale_y, 1.0) : 1.0;
            max_slot = std::max(max_slot, slot);
            max_iw = std::max(max_iw, iw);
            ++n_fused;
        } else {
            if (p.flags & B200COMP_REPLACE)
                return fail(B200COMP_EINVAL, "plan_create: B200COMP_REPLACE needs a placement the fused kernel resamples");
            // pre-resample with the generic kernels; the tile kernel then sees an identity-size overlay
            b200comp_plan::Pre pr;
            pr.src = p.src; pr.sp = p.src_pitch; pr.sw = p.sw; pr.sh = p.sh; pr.w = p.w; pr.h = p.h;
            pr.flags = vertical_first ? B200COMP_VERTICAL_FIRST : 0;
            if (need_h) pr.tx = ts.want_legacy(p.sw, p.w);
            if (need_v) pr.ty = ts.want_legacy(p.sh, p.h);
            pr.dst = nullptr; pr.scratch = nullptr;
            pre_index[i] = (int)plan->pre.size();
            plan->pre.push_back(pr);
            d.mode = 0;
            d.sw = p.w; d.sh = p.h;
            d.src_pitch = (p.w * 4 + 15) & ~15;
        }
    }

    stamp("placements classified");
    // ---- canvases / tiles ----
    std::vector<DevCanvas> hc((size_t)n_canvases);
    int64_t tiles = 0, algo = 0, step_records = 0;
    for (int c = 0; c < n_canvases; ++c) {
        const b200comp_canvas &cv = canvases[c];
        if (!cv.out || cv.W < 1 || cv.H < 1) return fail(B200COMP_EINVAL, "plan_create: canvas " + std::to_string(c) + " has no output or empty size");
        if (!aligned4(cv.out, cv.out_pitch) || cv.out_pitch < (int64_t)cv.W * 4 || (cv.bg && (!aligned4(cv.bg, cv.bg_pitch) || cv.bg_pitch < (int64_t)cv.W * 4)))
            return fail(B200COMP_EINVAL, "plan_create: canvas " + std::to_string(c) + " misaligned or bad pitch");
        if (cv.n_placements < 0 || cv.first_placement < 0 || (int64_t)cv.first_placement + cv.n_placements > n_placements)
            return fail(B200COMP_EINVAL, "plan_create: canvas " + std::to_string(c) + " placement range out of bounds");
        DevCanvas &d = hc[c];
        std::memset(&d, 0, sizeof d);
        d.out = cv.out; d.bg = cv.bg; d.out_pitch = cv.out_pitch; d.bg_pitch = cv.bg_pitch;
        d.solid = cv.solid_rgba; d.W = cv.W; d.H = cv.H;
        d.first = cv.first_placement; d.count = cv.n_placements;
        d.tiles_x = (cv.W + kTileW - 1) / kTileW;
        d.tiles_y = (cv.H + kTileH - 1) / kTileH;
        d.tile_base = tiles;
        plan->tiles_before.push_back(tiles);
        plan->records_before.push_back(tiles + step_records);
        tiles += (int64_t)d.tiles_x * d.tiles_y;
        if ((int64_t)d.tiles_x * d.tiles_y > INT32_MAX) return fail(B200COMP_EINVAL, "plan_create: canvas too large");
        plan->max_tiles = std::max(plan->max_tiles, d.tiles_x * d.tiles_y);
        algo += (int64_t)cv.W * cv.H * 4 * (cv.bg ? 2 : 1);
        for (int i = 0; i < cv.n_placements; ++i) {
            const b200comp_placement &p = placements[cv.first_placement + i];
            algo += (int64_t)p.sw * p.sh * 4;
            // tiles the box touches = records the binning pass writes for it
            const int64_t x0 = std::max<int64_t>(0, p.x), y0 = std::max<int64_t>(0, p.y);
            const int64_t x1 = std::min<int64_t>(cv.W, (int64_t)p.x + p.w), y1 = std::min<int64_t>(cv.H, (int64_t)p.y + p.h);
            if (x0 < x1 && y0 < y1)
                step_records += ((x1 - 1) / kTileW - x0 / kTileW + 1) * ((y1 - 1) / kTileH - y0 / kTileH + 1);
        }
    }
    plan->tiles_before.push_back(tiles);
    plan->records_before.push_back(tiles + step_records);
    plan->n_tiles = tiles;
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device);
        plan->G = std::max(1, std::min(1024, kCtasPerSm * sms));
        plan->stream_capacity = tiles + step_records + plan->G;
    }

    stamp("canvases / tiles");
    // ---- build tables on the host threads, upload everything ----
    // B200COMP_HOST_TABLES=1 builds the packed tables with libm on the host threads instead (validation)
    static const bool host_tables = std::getenv("B200COMP_HOST_TABLES") != nullptr;
    bool tables_on_device = !host_tables;
    const size_t tbytes = ((size_t)ts.total + 4) * sizeof(int32_t);
    auto dev_alloc = [&](void **p, size_t bytes) -> cudaError_t {
        cudaError_t e = lib_malloc_async(p, std::max<size_t>(bytes, 16), st);
        if (e == cudaSuccess) plan->owned.push_back(*p);
        return e;
    };
    CUDA_TRY(dev_alloc((void **)&plan->d_tables, tbytes));
    CUDA_TRY(dev_alloc((void **)&plan->d_placements, hp.size() * sizeof(DevPlacementT)));
    CUDA_TRY(dev_alloc((void **)&plan->d_canvases, hc.size() * sizeof(DevCanvas)));
    CUDA_TRY(dev_alloc((void **)&plan->d_status, sizeof(int)));
    for (auto &pr : plan->pre) {
        CUDA_TRY(dev_alloc((void **)&pr.dst, (size_t)((pr.w * 4 + 15) & ~15) * pr.h));
        if (pr.w != pr.sw && pr.h != pr.sh)
            CUDA_TRY(dev_alloc((void **)&pr.scratch, (size_t)std::max((int64_t)pr.sh * pr.w, (int64_t)pr.h * pr.sw) * 4));
    }
    for (int i = 0; i < n_placements; ++i) {
        DevPlacementT &d = hp[i];
        if (d.mode == 1) {
            d.plx = reinterpret_cast<const uint32_t *>(plan->d_tables + tref[i].first.b_off);
            d.ply = reinterpret_cast<const uint32_t *>(plan->d_tables + tref[i].second.b_off);
        } else if (pre_index[i] >= 0) {
            d.src = plan->pre[(size_t)pre_index[i]].dst;
        }
    }
    stamp("allocations");
    // ---- prepared cutouts (premultiplied, planar) + one TMA descriptor per resampled placement ----
    {
        typedef std::tuple<const uint8_t *, int, int, int64_t> SrcKey;
        std::map<SrcKey, int> prep_index;
        std::vector<PrepDesc> hprep;
        std::vector<CUtensorMap> hmaps;
        std::vector<int> map_of((size_t)std::max(1, n_placements), -1);
        std::vector<int> prep_of((size_t)std::max(1, n_placements), -1);
        std::vector<int64_t> flag_off;
        std::vector<size_t> prep_off;
        int64_t flag_words = 0;
        size_t prep_bytes = 0;
        struct PendingMap { int placement, prep; };
        std::vector<PendingMap> pending;
        int64_t max_words = 1;
        for (int i = 0; i < n_placements; ++i) {
            if (hp[i].mode != 1) continue;
            const b200comp_placement &p = placements[i];
            SrcKey key(p.src, p.sw, p.sh, p.src_pitch);
            auto it = prep_index.find(key);
            if (it == prep_index.end()) {
                PrepDesc pd;
                std::memset(&pd, 0, sizeof pd);
                pd.src = p.src;
                pd.src_pitch = p.src_pitch;
                pd.sw = p.sw;
                pd.sh = p.sh;
                pd.w4 = (p.sw + 3) / 4;
                pd.wq = (pd.w4 + 3) / 4;
                pd.vec_ok = ((reinterpret_cast<uintptr_t>(p.src) & 15u) == 0 && (p.src_pitch & 15) == 0) ? 1 : 0;
                prep_off.push_back(prep_bytes);
                prep_bytes += ((size_t)pd.w4 * 64 * ((p.sh + 3) / 4) + 255) & ~(size_t)255;
                flag_off.push_back(flag_words);
                flag_words += (int64_t)((p.sh + 3) / 4) * pd.wq;
                max_words = std::max<int64_t>(max_words, (int64_t)pd.w4 * ((p.sh + 3) / 4));
                it = prep_index.emplace(key, (int)hprep.size()).first;
                hprep.push_back(pd);
            }
            pending.push_back(PendingMap{i, it->second});
        }
        // overlays composited as they are (identity size, or pre-resampled into a plan-owned temporary) only need
        // the alpha summary: it lets the binning pass drop transparent tile steps and use opaque ones as occluders
        std::map<SrcKey, int> flags_only_index;
        std::vector<int> flags_only_of((size_t)std::max(1, n_placements), -1);
        for (int i = 0; i < n_placements; ++i) {
            if (hp[i].mode != 0) continue;
            SrcKey key(hp[i].src, hp[i].sw, hp[i].sh, (int64_t)hp[i].src_pitch);
            auto it = flags_only_index.find(key);
            if (it == flags_only_index.end()) {
                PrepDesc pd;
                std::memset(&pd, 0, sizeof pd);
                pd.src = hp[i].src;
                pd.src_pitch = hp[i].src_pitch;
                pd.sw = hp[i].sw;
                pd.sh = hp[i].sh;
                pd.w4 = (pd.sw + 3) / 4;
                pd.wq = (pd.w4 + 3) / 4;
                pd.vec_ok = ((reinterpret_cast<uintptr_t>(pd.src) & 15u) == 0 && (pd.src_pitch & 15) == 0) ? 1 : 0;
                prep_off.push_back(0);  // no prepared copy: dst stays null
                flag_off.push_back(flag_words);
                flag_words += (int64_t)((pd.sh + 3) / 4) * pd.wq;
                max_words = std::max<int64_t>(max_words, (int64_t)pd.w4 * ((pd.sh + 3) / 4));
                it = flags_only_index.emplace(key, (int)hprep.size()).first;
                hprep.push_back(pd);
            }
            flags_only_of[i] = it->second;
        }
        const size_t n_full_prep = prep_index.size();
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return fail(B200COMP_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        const cuuint32_t estride[3] = {1, 1, 1};
        if (!hprep.empty()) {
            uint8_t *d_prepared = nullptr;  // one allocation for every prepared cutout of the plan
            CUDA_TRY(dev_alloc((void **)&d_prepared, std::max<size_t>(prep_bytes, 16)));
            for (size_t j = 0; j < n_full_prep; ++j)  // the descriptors after these only produce the alpha summary
                hprep[j].dst = reinterpret_cast<uint32_t *>(d_prepared + prep_off[j]);
            // one tensor map per (prepared cutout, patch width class): dims (4 rows x word columns, 4 planes, row quads),
            // box = (4 * pwc, 4, kChunkQuads) -- every placement of that cutout in that class shares it
            std::map<std::pair<int, int>, int> class_map;
            for (const PendingMap &pm : pending) {
                const int i = pm.placement;
                const PrepDesc &pd = hprep[(size_t)pm.prep];
                auto it = class_map.find(std::make_pair(pm.prep, hp[i].pwc));
                if (it == class_map.end()) {
                    CUtensorMap tm;
                    const cuuint64_t gdim[3] = {(cuuint64_t)pd.w4 * 4, 4, (cuuint64_t)((pd.sh + 3) / 4)};
                    const cuuint64_t gstride[2] = {(cuuint64_t)pd.w4 * 16, (cuuint64_t)pd.w4 * 64};
                    const cuuint32_t box[3] = {(cuuint32_t)hp[i].pwc * 4, 4, (cuuint32_t)kChunkQuads};
                    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, pd.dst, gdim, gstride, box, estride,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS)
                        return fail(B200COMP_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for placement " + std::to_string(i));
                    it = class_map.emplace(std::make_pair(pm.prep, hp[i].pwc), (int)hmaps.size()).first;
                    hmaps.push_back(tm);
                }
                map_of[i] = it->second;
                prep_of[i] = pm.prep;
            }
            CUDA_TRY(dev_alloc((void **)&plan->d_prep, hprep.size() * sizeof(PrepDesc)));
            CUDA_TRY(dev_alloc((void **)&plan->d_flags, (size_t)flag_words * 4));
            plan->flag_bytes = (size_t)flag_words * 4;
            for (size_t j = 0; j < hprep.size(); ++j) hprep[j].flags = plan->d_flags + flag_off[j];
            CUDA_TRY(cudaMemcpyAsync(plan->d_prep, hprep.data(), hprep.size() * sizeof(PrepDesc), cudaMemcpyHostToDevice, st));
            plan->n_prep = (int)hprep.size();
            // grid.y = cutouts; grid.x = enough blocks that a plan with one large cutout (the stand-alone resize) still
            // fills the GPU, and not more than one block per 256 items of the largest cutout
            plan->prep_blocks_x = (int)std::max<int64_t>(1, std::min<int64_t>((max_words + 255) / 256,
                                                                              std::max<int64_t>(64, 148 * 16 / (int64_t)hprep.size())));
        }
        // 2-D maps: u32 pixels x rows.  Overlays composited as they are: 68 x 16-pixel boxes (one chunk each).
        // Canvases: 32-pixel x kTileH-row boxes with the 128-byte swizzle (tile_kernel.cuh ct_off).  Buffers TMA cannot
        // address (base not 16-byte aligned, pitch not a multiple of 16) keep a null map: generic loads/stores.
        auto encode_2d = [&](const void *base, int64_t pitch, int w, int h, int bw, int bh, bool swizzle, CUtensorMap *tm) -> bool {
            if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (pitch & 15) != 0 || pitch <= 0) return false;
            const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
            const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
            const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
            return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), gdim, gstride, box, estride,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        {
            std::map<SrcKey, int> overlay_map;
            for (int i = 0; i < n_placements; ++i) {
                if (hp[i].mode != 0) continue;
                SrcKey key(hp[i].src, hp[i].sw, hp[i].sh, (int64_t)hp[i].src_pitch);
                auto it = overlay_map.find(key);
                if (it == overlay_map.end()) {
                    CUtensorMap tm;
                    int idx = -1;
                    if (encode_2d(hp[i].src, hp[i].src_pitch, hp[i].sw, hp[i].sh, kOverlayBoxW, kIdentRows, false, &tm)) {
                        idx = (int)hmaps.size();
                        hmaps.push_back(tm);
                    }
                    it = overlay_map.emplace(key, idx).first;
                }
                map_of[i] = it->second;
            }
        }
        std::vector<int> bg_map_of((size_t)n_canvases, -1), out_map_of((size_t)n_canvases, -1);
        for (int c = 0; c < n_canvases; ++c) {
            CUtensorMap tm;
            if (hc[c].bg && encode_2d(hc[c].bg, hc[c].bg_pitch, hc[c].W, hc[c].H, 32, kTileH, true, &tm)) {
                bg_map_of[c] = (int)hmaps.size();
                hmaps.push_back(tm);
            }
            if (encode_2d(hc[c].out, hc[c].out_pitch, hc[c].W, hc[c].H, 32, kTileH, true, &tm)) {
                out_map_of[c] = (int)hmaps.size();
                hmaps.push_back(tm);
            }
        }
        CUDA_TRY(dev_alloc((void **)&plan->d_maps, std::max<size_t>(1, hmaps.size()) * sizeof(CUtensorMap)));
        if (!hmaps.empty())
            CUDA_TRY(cudaMemcpyAsync(plan->d_maps, hmaps.data(), hmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
        stamp("tensor maps encoded");
        CUDA_TRY(cudaStreamSynchronize(st));  // hmaps / hprep are locals
        stamp("maps uploaded (sync)");
        for (int i = 0; i < n_placements; ++i) {
            if (map_of[i] >= 0) hp[i].tmap = plan->d_maps + (size_t)map_of[i] * sizeof(CUtensorMap);
            const int pi = prep_of[i] >= 0 ? prep_of[i] : flags_only_of[i];
            if (pi >= 0) {
                const PrepDesc &pd = hprep[(size_t)pi];
                hp[i].flags = pd.flags;
                hp[i].wq = pd.wq;
                hp[i].sh4 = (pd.sh + 3) / 4;
            }
        }
        for (int c = 0; c < n_canvases; ++c) {
            hc[c].bg_map = bg_map_of[c] >= 0 ? plan->d_maps + (size_t)bg_map_of[c] * sizeof(CUtensorMap) : nullptr;
            hc[c].out_map = out_map_of[c] >= 0 ? plan->d_maps + (size_t)out_map_of[c] * sizeof(CUtensorMap) : nullptr;
        }
    }
    // command streams: records + ring read-ahead slack; per-tile counts; stream offsets
    {
        const int64_t K = (tiles + plan->G - 1) / plan->G;
        // every wave of a run has its own END records and read-ahead slack, its own rows of counts and offsets
        CUDA_TRY(dev_alloc((void **)&plan->d_streams,
                           (size_t)(plan->stream_capacity + (int64_t)kMaxWaves * (plan->G + kCmdBlk)) * sizeof(Cmd)));
        CUDA_TRY(dev_alloc((void **)&plan->d_bin, (size_t)plan->G * (size_t)(std::max<int64_t>(1, K) + kMaxWaves) * sizeof(int32_t)));
        CUDA_TRY(dev_alloc((void **)&plan->d_stream_off, (size_t)kMaxWaves * (plan->G + 2) * sizeof(int64_t)));
        CUDA_TRY(dev_alloc((void **)&plan->d_stream_len, (size_t)kMaxWaves * plan->G * sizeof(int64_t)));
        CUDA_TRY(dev_alloc((void **)&plan->d_dbg, 16 * sizeof(uint32_t)));
        CUDA_TRY(cudaMemsetAsync(plan->d_dbg, 0, 16 * sizeof(uint32_t), st));
        int max_count = 1;
        for (int c = 0; c < n_canvases; ++c) max_count = std::max(max_count, canvases[c].n_placements);
        plan->mask_chunks = (max_count + 31) / 32;
        CUDA_TRY(dev_alloc((void **)&plan->d_masks, (size_t)std::max<int64_t>(1, tiles) * plan->mask_chunks * 2 * sizeof(uint32_t)));
    }
    stamp("stream buffers allocated");
    int64_t n_fixed = 0;
    if (tables_on_device) {
        if (ts.has_legacy()) {  // legacy tables of the generic kernels: host-built, uploaded with the (still empty) packed regions
            ts.build(n_host_threads, true);
            CUDA_TRY(cudaMemcpyAsync(plan->d_tables, ts.host.data(), tbytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        std::string err;
        const int r = build_packed_on_device(ts, plan->d_tables, st, &n_fixed, &err);
        if (r < 0) return fail(B200COMP_ECUDA, err);
        if (r > 0) tables_on_device = false;
    }
    if (!tables_on_device) {
        ts.build(n_host_threads);
        CUDA_TRY(cudaMemcpyAsync(plan->d_tables, ts.host.data(), tbytes, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(plan->d_placements, hp.data(), hp.size() * sizeof(DevPlacementT), cudaMemcpyHostToDevice, st));
    std::vector<int4> hboxes(hp.size());
    for (size_t i = 0; i < hp.size(); ++i) hboxes[i] = make_int4(hp[i].x, hp[i].y, hp[i].w, hp[i].h);
    CUDA_TRY(dev_alloc((void **)&plan->d_boxes, hboxes.size() * sizeof(int4)));
    CUDA_TRY(cudaMemcpyAsync(plan->d_boxes, hboxes.data(), hboxes.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(plan->d_canvases, hc.data(), hc.size() * sizeof(DevCanvas), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(plan->d_status, 0, sizeof(int), st));
    stamp("tables built");
    CUDA_TRY(cudaStreamSynchronize(st));  // host staging vectors die with this scope
    stamp("descriptors uploaded (sync)");

    plan->slot_words = (max_slot + 31) & ~31;  // slots stay 128-byte aligned
    plan->iw_words = max_iw;
    CUDA_TRY(cudaFuncSetAttribute(composite_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemBytes));
    {
        // deepest ring that keeps the CTAs per SM the kernel is built for (asked of the occupancy calculator, so the
        // reserved shared memory and register limits of the device at hand are respected); B200COMP_RING=n overrides
        const char *env = std::getenv("B200COMP_RING");
        const int forced = env && env[0] ? std::max(kPRingMin, std::min(kPRingMax, std::atoi(env))) : 0;
        plan->n_ring = kPRingMin;
        for (int n = forced ? forced : kPRingMax; n > kPRingMin; --n) {
            const size_t bytes = tile_smem_bytes(plan->slot_words, plan->iw_words, n);
            int resident = 0;
            if (bytes <= kMaxSmemBytes &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, composite_slab_kernel, kThreads, bytes) == cudaSuccess &&
                (resident >= kCtasPerSm || forced)) {
                plan->n_ring = n;
                break;
            }
        }
        cudaGetLastError();
    }
    plan->smem_bytes = tile_smem_bytes(plan->slot_words, plan->iw_words, plan->n_ring);
    if (plan->smem_bytes > kMaxSmemBytes) return fail(B200COMP_EINTERNAL, "plan_create: tile kernel shared memory over the limit");

    plan->info[B200COMP_INFO_ALGORITHMIC_BYTES] = algo;
    // 3 binning kernels + tile kernel per wave of a whole-plan run
    plan->info[B200COMP_INFO_LAUNCHES_PER_RUN] = 4 * wave_count(plan, tiles, n_canvases) + (plan->n_prep > 0 ? 1 : 0);
    for (auto &pr : plan->pre) plan->info[B200COMP_INFO_LAUNCHES_PER_RUN] += (pr.w != pr.sw && pr.h != pr.sh) ? 2 : 1;
    plan->info[B200COMP_INFO_FUSED_PLACEMENTS] = n_fused;
    plan->info[B200COMP_INFO_IDENTITY_PLACEMENTS] = n_ident;
    plan->info[B200COMP_INFO_PRERESAMPLED_PLACEMENTS] = (int64_t)plan->pre.size();
    plan->info[B200COMP_INFO_COEFF_BYTES] = (int64_t)tbytes;
    plan->info[B200COMP_INFO_SMEM_BYTES] = (int64_t)plan->smem_bytes;
    plan->info[B200COMP_INFO_TILES] = tiles;

    guard.p = nullptr;
    *out_plan = plan;
    return 0;
}

// Stage 1 of a run: everything the tile kernel reads besides the caller's buffers (pre-resampled
// overlays for extreme scales, prepared cutouts + alpha summaries).
int b200comp_plan_prepare(b200comp_plan *plan, void *stream) {
    if (!plan) return fail(B200COMP_EINVAL, "plan_prepare: null plan");
    cudaStream_t st = S(stream);
    if (plan->profile) {
        for (auto &e : plan->prof_prepare) if (!e) CUDA_TRY(cudaEventCreate(&e));
        CUDA_TRY(cudaEventRecord(plan->prof_prepare[0], st));
    }
    for (auto &pr : plan->pre) {
        const int32_t *t = plan->d_tables;
        int rc = resample_two_pass(pr.src, pr.sw, pr.sh, pr.sp, pr.dst, pr.w, pr.h, (int64_t)((pr.w * 4 + 15) & ~15), t + pr.tx.k_off,
                                   t + pr.tx.b_off, pr.tx.ks, t + pr.ty.k_off, t + pr.ty.b_off, pr.ty.ks, pr.scratch,
                                   pr.flags, st);
        if (rc) return rc;
    }
    if (plan->n_prep > 0) {
        // premultiplied planar copies of the cutouts the tile kernel resamples (re-made every run, so
        // the plan never shows stale pixels if the caller rewrites a cutout between runs)
        CUDA_TRY(cudaMemsetAsync(plan->d_flags, 0, plan->flag_bytes, st));
        for (int p0 = 0; p0 < plan->n_prep; p0 += 65535) {
            const int np = std::min(65535, plan->n_prep - p0);
            prepare_cutouts_kernel<<<dim3((unsigned)plan->prep_blocks_x, (unsigned)np), 256, 0, st>>>(plan->d_prep + p0);
        }
        CUDA_TRY(cudaGetLastError());
    }
    if (plan->profile) {
        CUDA_TRY(cudaEventRecord(plan->prof_prepare[1], st));
        plan->prof_prepare_valid = true;
    }
    return 0;
}

// Stage 2: binning + the fused tile kernel over canvases [first, first + count).
//
// A large run is cut into waves of whole canvases.  Every wave has its own slice of the plan's binning buffers
// (masks by absolute tile, count rows, stream offsets / lengths / cursor, a region of the record array), so the
// waves only depend on the prepared cutouts.  Binning of all waves is queued on the plan's `s_bin`; the tile kernel of
// wave w waits for its binning and runs on the caller's stream (even waves) or on `s_tile` (odd waves).  The binning
// kernels are small blocks that fit beside the two resident tile CTAs of an SM, and most of their time is the HBM copy
// of the tiles nothing is drawn on, which the issue-bound tile kernel leaves room for; the next wave's tile CTAs
// fill the slots the previous wave's stragglers free.  Everything is joined back into the caller's stream.
int b200comp_plan_run_canvases(b200comp_plan *plan, int first, int count, void *stream) {
    if (!plan) return fail(B200COMP_EINVAL, "plan_run_canvases: null plan");
    if (first < 0 || count < 0 || (int64_t)first + count > plan->n_canvases)
        return fail(B200COMP_EINVAL, "plan_run_canvases: canvas range out of bounds");
    cudaStream_t st = S(stream);
    if (count == 0) return 0;
    // Runs of one plan share its command-stream buffers: they must be ordered on the stream.
    const int64_t run_tile0 = plan->tiles_before[(size_t)first];
    const int64_t run_tiles = plan->tiles_before[(size_t)first + count] - run_tile0;
    if (run_tiles >= (int64_t)1 << 31) return fail(B200COMP_EINVAL, "plan_run_canvases: more than 2^31 tiles in one run");
    const int G = plan->G;
    const unsigned gx = (unsigned)((plan->max_tiles + kBinWarps - 1) / kBinWarps);  // warp = tile
    cudaEvent_t pe[3] = {nullptr, nullptr, nullptr};
    if (plan->profile) {
        for (auto &e : pe) {
            CUDA_TRY(cudaEventCreate(&e));
            plan->prof_events.push_back(e);
        }
        CUDA_TRY(cudaEventRecord(pe[0], st));
    }
    // B200COMP_DEBUG_SYNC=1: one wave, synchronise after every launch so a device fault names its kernel
    static const bool debug_sync = std::getenv("B200COMP_DEBUG_SYNC") != nullptr;
    auto checkpoint = [&](const char *what) -> int {
        if (!debug_sync) return 0;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return fail(B200COMP_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return 0;
    };
    // B200COMP_NO_CULL=1 keeps the steps an opaque, tile-covering later placement hides (A/B parity tests)
    const char *no_cull = std::getenv("B200COMP_NO_CULL");
    const int cull = !(no_cull && no_cull[0] == '1');

    const int n_waves = debug_sync ? 1 : wave_count(plan, run_tiles, count);
    plan->last_waves = n_waves;
    if (n_waves > 1) {
        if (!plan->s_bin) CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_bin, cudaStreamNonBlocking));
        if (!plan->s_tile) CUDA_TRY(cudaStreamCreateWithFlags(&plan->s_tile, cudaStreamNonBlocking));
        if (!plan->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
        if (!plan->ev_tile) CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_tile, cudaEventDisableTiming));
        for (int w = 0; w < n_waves; ++w)
            if (!plan->ev_bin[w]) CUDA_TRY(cudaEventCreateWithFlags(&plan->ev_bin[w], cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(plan->ev_fork, st));  // behind the prepare pass and every earlier run
        CUDA_TRY(cudaStreamWaitEvent(plan->s_bin, plan->ev_fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(plan->s_tile, plan->ev_fork, 0));
    }
    int64_t bin_row0 = 0;  // first count row (of G entries each... rows are [G][K_wave]) of this wave in d_bin
    int c0 = first;
    bool odd_used = false;
    for (int w = 0; w < n_waves; ++w) {
        // canvases of this wave: cut at the canvas whose first tile is nearest to an equal share of the tiles
        int c1 = first + count;
        if (w + 1 < n_waves) {
            const int64_t target = run_tile0 + run_tiles * (w + 1) / n_waves;
            c1 = (int)(std::lower_bound(plan->tiles_before.begin() + c0 + 1, plan->tiles_before.begin() + first + count, target) -
                       plan->tiles_before.begin());
            c1 = std::max(c0 + 1, std::min(c1, first + count - (n_waves - 1 - w)));
        }
        const int64_t tile0 = plan->tiles_before[(size_t)c0];
        const int64_t n_tiles = plan->tiles_before[(size_t)c1] - tile0;
        const int K = (int)((n_tiles + G - 1) / G);
        int32_t *bin = plan->d_bin + bin_row0 * G;
        bin_row0 += K;
        int64_t *stream_off = plan->d_stream_off + (size_t)w * (G + 2);
        int64_t *stream_len = plan->d_stream_len + (size_t)w * G;
        unsigned long long *cursor = reinterpret_cast<unsigned long long *>(stream_off + G);  // record allocator
        const int64_t rec0 = plan->records_before[(size_t)c0] - plan->records_before[(size_t)first] + (int64_t)w * (G + kCmdBlk);
        const int64_t capacity = plan->records_before[(size_t)c1] - plan->records_before[(size_t)c0] + G;
        Cmd *streams = plan->d_streams + rec0;
        uint32_t *masks = plan->d_masks + (tile0 - run_tile0) * (int64_t)plan->mask_chunks * 2;
        cudaStream_t sb = n_waves > 1 ? plan->s_bin : st;
        for (int b0 = c0; b0 < c1; b0 += 65535) {  // grid.y is limited to 65535 canvases per launch
            const int nc = std::min(65535, c1 - b0);
            bin_count_kernel<<<dim3(gx, (unsigned)nc), kBinWarps * 32, 0, sb>>>(
                plan->d_canvases + b0, plan->d_placements, plan->d_boxes, tile0, G, K, bin, masks, plan->mask_chunks,
                plan->iw_words, b0 == c0 ? cursor : nullptr, plan->d_status, cull);
        }
        if (int rc = checkpoint("bin_count_kernel")) return rc;
        bin_scan_kernel<<<(unsigned)((G + 7) / 8), 256, 0, sb>>>(bin, G, K, n_tiles, stream_off, stream_len, cursor, streams,
                                                                 capacity, plan->d_status);
        if (int rc = checkpoint("bin_scan_kernel")) return rc;
        for (int b0 = c0; b0 < c1; b0 += 65535) {
            const int nc = std::min(65535, c1 - b0);
            bin_fill_kernel<<<dim3(gx, (unsigned)nc), kBinWarps * 32, 0, sb>>>(
                plan->d_canvases + b0, b0, plan->d_placements, tile0, G, K, bin, masks, plan->mask_chunks, stream_off, streams,
                capacity, plan->d_maps, reinterpret_cast<const uint32_t *>(plan->d_tables));
        }
        if (int rc = checkpoint("bin_fill_kernel")) return rc;
        cudaStream_t stt = st;
        if (n_waves > 1) {
            CUDA_TRY(cudaEventRecord(plan->ev_bin[w], sb));
            if (w & 1) {
                stt = plan->s_tile;
                odd_used = true;
            }
            CUDA_TRY(cudaStreamWaitEvent(stt, plan->ev_bin[w], 0));
        }
        if (pe[1]) CUDA_TRY(cudaEventRecord(pe[1], st));
        static const size_t extra_smem = [] {  // B200COMP_EXTRA_SMEM=bytes: occupancy experiments (1 CTA per SM)
            const char *e = std::getenv("B200COMP_EXTRA_SMEM");
            return e ? (size_t)std::atol(e) : (size_t)0;
        }();
        composite_slab_kernel<<<(unsigned)G, kThreads, plan->smem_bytes + extra_smem, stt>>>(
            streams, stream_off, stream_len, plan->d_canvases, plan->d_maps, reinterpret_cast<const uint32_t *>(plan->d_tables),
            plan->slot_words, plan->iw_words, plan->n_ring, plan->d_status, plan->d_dbg);
        CUDA_TRY(cudaGetLastError());
        c0 = c1;
    }
    if (odd_used) {  // join: the caller's stream carries the even waves and waits for the odd ones
        CUDA_TRY(cudaEventRecord(plan->ev_tile, plan->s_tile));
        CUDA_TRY(cudaStreamWaitEvent(st, plan->ev_tile, 0));
    }
    if (pe[2]) {
        CUDA_TRY(cudaEventRecord(pe[2], st));
        ++plan->prof_runs;
    }
    if (int rc = checkpoint("composite_slab_kernel")) return rc;
    return 0;
}

int b200comp_plan_run(b200comp_plan *plan, void *stream) {
    int rc = b200comp_plan_prepare(plan, stream);
    if (rc) return rc;
    return b200comp_plan_run_canvases(plan, 0, plan->n_canvases, stream);
}

// internal (tools/): copy the command streams of the last run to the host.  out: records of 16 words;
// offs: G + 1 stream offsets.  Returns the number of records copied or a negative error.
int64_t b200comp_plan_debug_streams_(b200comp_plan *plan, uint32_t *out, int64_t max_records, int64_t *offs, int *n_streams) {
    if (!plan || !out || !offs) return B200COMP_EINVAL;
    cudaDeviceSynchronize();
    cudaMemcpy(offs, plan->d_stream_off, (size_t)(plan->G + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost);
    if (n_streams) *n_streams = plan->G;
    const int64_t n = std::min<int64_t>(max_records, std::min<int64_t>(offs[plan->G], plan->stream_capacity));
    cudaMemcpy(out, plan->d_streams, (size_t)n * sizeof(Cmd), cudaMemcpyDeviceToHost);
    return n;
}

// internal (tests/): packed coefficient tables of the given (in, out) size pairs built on the device (with the
// libm fix-up) and on the host; returns the number of differing words (0 = bit-identical) or a negative error.
int64_t b200comp_debug_compare_tables_(const int *in_sizes, const int *out_sizes, int n, int64_t *n_fixed) {
    TableSet ts;
    for (int i = 0; i < n; ++i) {
        if (in_sizes[i] < 1 || out_sizes[i] < 1) return B200COMP_EINVAL;
        if (in_sizes[i] == out_sizes[i]) { ts.want_packed(in_sizes[i], out_sizes[i], true); continue; }
        if (packed_words(lanczos_ksize(in_sizes[i], out_sizes[i])) == 0) continue;  // not a fused-path table
        ts.want_packed(in_sizes[i], out_sizes[i], false);
    }
    const size_t tbytes = ((size_t)ts.total + 4) * sizeof(int32_t);
    int32_t *d = nullptr;
    if (cudaMalloc((void **)&d, tbytes) != cudaSuccess) return B200COMP_ENOMEM;
    cudaMemset(d, 0xff, tbytes);
    std::string err;
    const int r = build_packed_on_device(ts, d, nullptr, n_fixed, &err);
    if (r != 0) {
        cudaFree(d);
        return r < 0 ? fail(B200COMP_ECUDA, err) : fail(B200COMP_EINTERNAL, "fix-up list overflow");
    }
    std::vector<int32_t> got((size_t)ts.total + 4);
    cudaMemcpy(got.data(), d, tbytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    ts.build(0);
    int64_t bad = 0;
    for (const TableSet::Key &key : ts.order) {
        const TableRef &ref = ts.refs.at(key);
        const int rw = coef_row_words(ref.ks);
        for (int64_t j = 0; j < key.out_size; ++j)
            for (int w = 0; w < 3 * ref.ks; ++w)
                bad += got[(size_t)(ref.b_off + j * rw + w)] != ts.host[(size_t)(ref.b_off + j * rw + w)];
    }
    return bad;
}

// internal (tools/): phase cycle counters of a -DB200COMP_PROFILE=1 build; reset after reading
int b200comp_debug_profile_(unsigned long long *out16) {
#if defined(B200COMP_PROFILE) && B200COMP_PROFILE
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_prof, sizeof(unsigned long long) * 16);
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_prof, z, sizeof z);
    return 0;
#else
    (void)out16;
    return B200COMP_EINVAL;
#endif
}

int b200comp_plan_profile(b200comp_plan *plan, int enable) {
    if (!plan) return fail(B200COMP_EINVAL, "plan_profile: null plan");
    plan->profile = enable != 0;
    return 0;
}

int b200comp_plan_profile_read(b200comp_plan *plan, double ms[3], int *runs) {
    if (!plan || !ms) return fail(B200COMP_EINVAL, "plan_profile_read: null argument");
    ms[0] = ms[1] = ms[2] = 0.0;
    if (runs) *runs = plan->prof_runs;
    if (!plan->prof_events.empty()) CUDA_TRY(cudaEventSynchronize(plan->prof_events.back()));
    for (size_t i = 0; i + 2 < plan->prof_events.size(); i += 3) {
        float a = 0.f, b = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&a, plan->prof_events[i], plan->prof_events[i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&b, plan->prof_events[i + 1], plan->prof_events[i + 2]));
        ms[1] += a;
        ms[2] += b;
    }
    // the prepare kernel runs once per b200comp_plan_run; only the last one is still bracketed
    if (plan->prof_prepare_valid && plan->prof_runs > 0) {
        float p = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&p, plan->prof_prepare[0], plan->prof_prepare[1]));
        ms[0] = (double)p * plan->prof_runs;
    }
    for (cudaEvent_t e : plan->prof_events) cudaEventDestroy(e);
    plan->prof_events.clear();
    plan->prof_runs = 0;
    plan->prof_prepare_valid = false;
    return 0;
}

int b200comp_plan_info(const b200comp_plan *plan, int64_t *info) {
    if (!plan || !info) return fail(B200COMP_EINVAL, "plan_info: null argument");
    std::memcpy(info, plan->info, sizeof plan->info);
    return 0;
}

// Reads the device status word (shared-memory sizing violations); synchronises the stream.
int b200comp_plan_check(b200comp_plan *plan, void *stream) {
    if (!plan) return fail(B200COMP_EINVAL, "plan_check: null plan");
    int h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, plan->d_status, sizeof h, cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    if (h & kStatusWatchdog) {
        uint32_t d[16] = {0};
        cudaMemcpy(d, plan->d_dbg, sizeof d, cudaMemcpyDeviceToHost);
        char msg[200];
        std::snprintf(msg, sizeof msg, "tile kernel watchdog: wait tag 0x%x parity %u aux %u in CTA %u thread %u never came true",
                      d[0], d[3], d[4], d[1], d[2]);
        return fail(B200COMP_EINTERNAL, msg);
    }
    if (h != 0) return fail(B200COMP_EINTERNAL, "binning reported a sizing violation (status " + std::to_string(h) + ")");
    return 0;
}

// internal (host_api.cu): copy the status word into pinned host memory in stream order, no synchronisation --
// the host-buffer pipeline reads it after the event that follows the copy
int b200comp_plan_status_async_(b200comp_plan *plan, int *pinned_host_dst, void *stream) {
    if (!plan || !pinned_host_dst) return fail(B200COMP_EINVAL, "plan_status_async: null argument");
    CUDA_TRY(cudaMemcpyAsync(pinned_host_dst, plan->d_status, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
    return 0;
}

int b200comp_plan_last_records(b200comp_plan *plan, void *stream, int64_t *records) {
    if (!plan || !records) return fail(B200COMP_EINVAL, "plan_last_records: null argument");
    // one cursor per wave of the last run, at the end of the wave's row of stream offsets
    std::vector<int64_t> rows((size_t)kMaxWaves * (plan->G + 2));
    CUDA_TRY(cudaMemcpyAsync(rows.data(), plan->d_stream_off, rows.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, S(stream)));
    CUDA_TRY(cudaStreamSynchronize(S(stream)));
    *records = 0;
    for (int w = 0; w < plan->last_waves; ++w) *records += rows[(size_t)w * (plan->G + 2) + plan->G];
    return 0;
}

int b200comp_composite_batch(const b200comp_canvas *canvases, int n_canvases, const b200comp_placement *placements,
                             int n_placements, void *stream) {
    b200comp_plan *plan = nullptr;
    int rc = b200comp_plan_create(canvases, n_canvases, placements, n_placements, 0, stream, &plan);
    if (rc) return rc;
    rc = b200comp_plan_run(plan, stream);
    if (!rc) rc = b200comp_plan_check(plan, stream);
    b200comp_plan_destroy(plan);
    return rc;
}

}  // extern "C"
#pragma GCC visibility pop
