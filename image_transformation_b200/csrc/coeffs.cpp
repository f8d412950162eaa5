// Host-side LANCZOS coefficient tables for the 8-bit separable resampler.
//
// Reproduces what `obj.resize((w, h), Image.LANCZOS)` (/root/reference/compositor.py:20)
// makes Pillow compute in Resample.c (precompute_coeffs + normalize_coeffs_8bpc):
// double precision, libm sin, truncating (int) conversions, 22-bit fixed point.
// Built WITHOUT fast-math / fp-contract so every double operation rounds once,
// exactly as in the scalar C the reference links against.  CUDA's sin() is not
// bit-identical to glibc's, which is why the tables are made on the host.
#include <cmath>
#include <cstdint>
#include <vector>

#include "coeffs.h"

namespace b200comp {

static const double kPi = 3.14159265358979323846;
static const int kPrecisionBits = 22;  // 32 - 8 - 2

static inline double sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * kPi;
    return std::sin(x) / x;
}

static inline double lanczos3(double x) {
    if (-3.0 <= x && x < 3.0) return sinc(x) * sinc(x / 3);
    return 0.0;
}

int lanczos_ksize(int in_size, int out_size) {
    double scale = static_cast<double>(in_size) / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    return static_cast<int>(std::ceil(3.0 * filterscale)) * 2 + 1;
}

// Taps of ONE output sample `xx` (the body of precompute_coeffs' loop + normalize_coeffs_8bpc).
// row: ksize int32 (zero padded past n).  `w` is scratch of ksize doubles.
static void lanczos_row_impl(int in_size, int out_size, int xx, double scale, double support, double inv_filterscale,
                             int ksize, double *w, int32_t *row, int *lo_out, int *n_out) {
    (void)out_size;
    const double center = 0.0 + (xx + 0.5) * scale;
    int lo = static_cast<int>(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = static_cast<int>(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    double total = 0.0;
    for (int x = 0; x < n; ++x) {
        const double v = lanczos3((x + lo - center + 0.5) * inv_filterscale);
        w[x] = v;
        total += v;
    }
    for (int x = 0; x < n; ++x) {
        double v = w[x];
        if (total != 0.0) v /= total;
        row[x] = v < 0 ? static_cast<int32_t>(-0.5 + v * (1 << kPrecisionBits))
                       : static_cast<int32_t>(0.5 + v * (1 << kPrecisionBits));
    }
    for (int x = n; x < ksize; ++x) row[x] = 0;
    *lo_out = lo;
    *n_out = n;
}

int build_lanczos_table(int in_size, int out_size, int32_t *k, int32_t *bounds) {
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale;
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    const double inv_filterscale = 1.0 / filterscale;
    std::vector<double> w(static_cast<size_t>(ksize));
    for (int xx = 0; xx < out_size; ++xx) {
        int lo, n;
        lanczos_row_impl(in_size, out_size, xx, scale, support, inv_filterscale, ksize, w.data(),
                         k + static_cast<size_t>(xx) * ksize, &lo, &n);
        bounds[2 * xx] = lo;
        bounds[2 * xx + 1] = n;
    }
    return ksize;
}

int build_lanczos_row(int in_size, int out_size, int xx, int32_t *row, int *lo, int *n) {
    const double scale = static_cast<double>(in_size) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale;
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    std::vector<double> w(static_cast<size_t>(ksize));
    lanczos_row_impl(in_size, out_size, xx, scale, support, 1.0 / filterscale, ksize, w.data(), row, lo, n);
    return ksize;
}

}  // namespace b200comp
