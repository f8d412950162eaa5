// Device-side construction of the packed LANCZOS coefficient planes.
//
// The reference's coefficients come from double-precision libm arithmetic on the host
// (Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc, reached from compositor.py:20).
// CUDA's sin() is accurate to 2 ulp but not bit-identical to glibc's, so this kernel does NOT decide
// borderline roundings: every IEEE-exact step (scale, centre, window bounds, the sinc arguments, the
// tap order of the normalising sum) is replayed with the host's operation order, the fixed-point
// value v = w * 2^22 is formed, and an output sample whose v lies within 1e-5 of a rounding boundary
// (x.5) for any tap is reported in a fix-up list; the host recomputes exactly those samples with libm
// and patches them (b200comp.cu).  The two sin() implementations differ by < 1e-8 in v (relative
// error ~3e-15 times |v| <= 4.2e6), three orders of magnitude below the 1e-5 guard band, so every
// sample NOT reported is bit-identical to the host table.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200comp {

struct CoefJob {
    int64_t planes_off;  // word offset of planes[3*nw][out] in the table buffer
    int32_t in_size, out_size;
    int32_t nw;          // words per output sample
    int32_t identity;    // skipped pass: one tap of 1.0 at the sample itself
};

struct CoefFix {
    int32_t job, j;
};

constexpr int kMaxTaps = 17;  // nw <= 5 -> at most 17 taps
constexpr double kPi = 3.14159265358979323846;

__device__ __forceinline__ double sinc_dev(double x) {
    if (x == 0.0) return 1.0;
    x = __dmul_rn(x, kPi);
    return __ddiv_rn(sin(x), x);
}
__device__ __forceinline__ double lanczos3_dev(double x) {
    if (-3.0 <= x && x < 3.0) return __dmul_rn(sinc_dev(x), sinc_dev(__ddiv_rn(x, 3.0)));
    return 0.0;
}

__global__ void __launch_bounds__(128) build_packed_tables_kernel(const CoefJob *__restrict__ jobs, int job0,
                                                                   uint32_t *__restrict__ tables,
                                                                   CoefFix *__restrict__ fix, int *__restrict__ fix_count,
                                                                   int fix_cap) {
    const CoefJob job = jobs[blockIdx.y];
    const int n_out = job.out_size, nw = job.nw;
    uint32_t *planes = tables + job.planes_off;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += gridDim.x * blockDim.x) {
        uint32_t pl[3][5];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int i = 0; i < 5; ++i) pl[p][i] = 0u;
        bool uncertain = false;
        if (job.identity) {
            // k = [1 << 22] at the sample itself: byte 2 of the tap is 64
            const int pos = j & 3;
            pl[2][0] = 64u << (8 * pos);
        } else {
            const double scale = __ddiv_rn((double)job.in_size, (double)n_out);
            const double filterscale = scale < 1.0 ? 1.0 : scale;
            const double support = __dmul_rn(3.0, filterscale);
            const double ss = __ddiv_rn(1.0, filterscale);
            const double center = __dmul_rn((double)j + 0.5, scale);
            int lo = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
            if (lo < 0) lo = 0;
            int hi = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
            if (hi > job.in_size) hi = job.in_size;
            const int n = min(hi - lo, kMaxTaps);
            double w[kMaxTaps];
            double total = 0.0;
#pragma unroll
            for (int t = 0; t < kMaxTaps; ++t) {
                if (t < n) {
                    const double x = __dmul_rn(__dadd_rn(__dsub_rn((double)(t + lo), center), 0.5), ss);
                    w[t] = lanczos3_dev(x);
                    total = __dadd_rn(total, w[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < kMaxTaps; ++t) {
                if (t < n) {
                    double v = w[t];
                    if (total != 0.0) v = __ddiv_rn(v, total);
                    const double f = __dmul_rn(v, 4194304.0);
                    const double r = v < 0 ? __dadd_rn(-0.5, f) : __dadd_rn(0.5, f);
                    const int k = (int)r;  // truncation, as the C cast
                    // distance of f from the nearest rounding boundary (integer + 0.5)
                    const double fr = f - floor(f);
                    if (fabs(fr - 0.5) < 1e-5) uncertain = true;
                    const int pos = (lo & 3) + t, word = pos >> 2, sh = 8 * (pos & 3);
#pragma unroll
                    for (int i = 0; i < 5; ++i)
                        if (i == word) {
                            pl[0][i] |= (uint32_t)(k & 0xff) << sh;
                            pl[1][i] |= (uint32_t)((k >> 8) & 0xff) << sh;
                            pl[2][i] |= (uint32_t)((k >> 16) & 0xff) << sh;
                        }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int i = 0; i < 5; ++i)
                if (i < nw) planes[(int64_t)j * coef_row_words(nw) + p * nw + i] = pl[p][i];
        if (uncertain) {
            const int idx = atomicAdd(fix_count, 1);
            if (idx < fix_cap) {
                fix[idx].job = job0 + (int)blockIdx.y;
                fix[idx].j = j;
            }
        }
    }
}

struct WordPatch {
    int64_t off;
    uint32_t value;
    uint32_t pad_;
};

__global__ void patch_words_kernel(uint32_t *__restrict__ tables, const WordPatch *__restrict__ patches, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tables[patches[i].off] = patches[i].value;
}

}  // namespace b200comp
