/*
 * compositor_oracle.c -- CPU restatement of the reference compositor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under image_transformation_b200/ may call,
 * link or import this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU
 * baseline, never as the product path.
 *
 * What is restated
 * ----------------
 * The reference hot path is /root/reference/compositor.py:6-22 (composite) and
 * /root/reference/background_resizing.py:11-33,36-98 (median colour, solid fill,
 * gradient fill).  Those 137 lines of Python delegate all arithmetic to two
 * third-party dependencies that are NOT vendored under /root/reference:
 *
 *   pillow==11.3.0 (requirements.txt:29; the oracle container has 12.2.0)
 *       src/libImaging/Resample.c       precompute_coeffs, normalize_coeffs_8bpc,
 *                                       ImagingResampleHorizontal_8bpc / Vertical_8bpc
 *       src/libImaging/Convert.c        rgbA2rgba (premultiply), rgba2rgbA (un-premultiply)
 *       src/libImaging/AlphaComposite.c ImagingAlphaComposite
 *       src/PIL/Image.py                Image.resize (identity short-cut, RGBA->RGBa
 *                                       round trip, 12.x tall-image vertical-first
 *                                       branch), Image.alpha_composite (crop/paste clip)
 *   numpy==2.3.3 (requirements.txt:23)  np.median (mean of the two middle order
 *                                       statistics, float64), float32 lerp
 *
 * Their published algorithms are restated below from scratch.  PARITY PINNING:
 * tests/golden/ holds input/output vectors produced by running the UNMODIFIED
 * reference modules (imported from /root/reference with the installed Pillow
 * 12.2.0 / NumPy 2.3.5) through tests/golden/make_golden.py; tests/test_oracle.py
 * checks every function in this file against them, plus the reference's own
 * known-answer test (tests/test_compositor.py:5-11).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PRECISION_BITS 22 /* Resample.c: 32 - 8 - 2 */

typedef struct {
    const uint8_t *src; /* cutout, RGBA interleaved, tightly packed sw*4 per row */
    int32_t sw, sh;
    int32_t x1, y1, x2, y2; /* box after int() truncation (compositor.py:16) */
} orc_placement;

/* ---- Convert.c arithmetic ---------------------------------------------- */

/* ImagingUtils.h MULDIV255 */
static inline uint32_t muldiv255(uint32_t a, uint32_t b) {
    uint32_t t = a * b + 128;
    return ((t >> 8) + t) >> 8;
}

/* rgbA2rgba: RGBA -> RGBa, used by Image.resize for RGBA (PIL Image.py:2406-2407) */
void orc_premultiply(const uint8_t *in, uint8_t *out, size_t npx) {
    for (size_t i = 0; i < npx; i++) {
        uint32_t a = in[4 * i + 3];
        out[4 * i + 0] = (uint8_t)muldiv255(in[4 * i + 0], a);
        out[4 * i + 1] = (uint8_t)muldiv255(in[4 * i + 1], a);
        out[4 * i + 2] = (uint8_t)muldiv255(in[4 * i + 2], a);
        out[4 * i + 3] = (uint8_t)a;
    }
}

/* rgba2rgbA: RGBa -> RGBA (PIL Image.py:2409): truncating divide, clip to 255 */
void orc_unpremultiply(const uint8_t *in, uint8_t *out, size_t npx) {
    for (size_t i = 0; i < npx; i++) {
        uint32_t a = in[4 * i + 3];
        if (a == 255 || a == 0) {
            out[4 * i + 0] = in[4 * i + 0];
            out[4 * i + 1] = in[4 * i + 1];
            out[4 * i + 2] = in[4 * i + 2];
        } else {
            for (int c = 0; c < 3; c++) {
                uint32_t v = (255u * in[4 * i + c]) / a;
                out[4 * i + c] = (uint8_t)(v > 255 ? 255 : v);
            }
        }
        out[4 * i + 3] = (uint8_t)a;
    }
}

/* ---- Resample.c: coefficients ------------------------------------------ */

static double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}

static double lanczos_filter(double x) {
    /* truncated sinc, support 3 */
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}

int orc_ksize(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = 3.0 * filterscale;
    return (int)ceil(support) * 2 + 1;
}

/* precompute_coeffs + normalize_coeffs_8bpc for the full-image box (0, in_size).
 * k: out_size*ksize int32 (zero padded beyond xmax), bounds: out_size*(xmin,xmax).
 * Returns ksize. */
int orc_coeffs(int in_size, int out_size, int32_t *k, int32_t *bounds) {
    double scale, filterscale, support;
    filterscale = scale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    support = 3.0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double *w = (double *)malloc(sizeof(double) * (size_t)ksize);
    if (!w) return -1;
    double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; xx++) {
        double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            double v = lanczos_filter((x + xmin - center + 0.5) * ss);
            w[x] = v;
            ww += v;
        }
        for (x = 0; x < xmax; x++)
            if (ww != 0.0) w[x] /= ww;
        int32_t *kk = k + (size_t)xx * ksize;
        for (x = 0; x < xmax; x++) {
            if (w[x] < 0)
                kk[x] = (int32_t)(-0.5 + w[x] * (1 << ORC_PRECISION_BITS));
            else
                kk[x] = (int32_t)(0.5 + w[x] * (1 << ORC_PRECISION_BITS));
        }
        for (; x < ksize; x++) kk[x] = 0;
        bounds[2 * xx + 0] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    free(w);
    return ksize;
}

/* ---- Resample.c: passes (8 bits per channel, 4 channels) --------------- */

static inline uint8_t clip8(int32_t v) {
    v >>= ORC_PRECISION_BITS; /* arithmetic shift, as clip8_lookups[in >> PRECISION_BITS] */
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* in: sh x sw, out: sh x w */
void orc_resample_h(const uint8_t *in, int sw, int sh, uint8_t *out, int w, const int32_t *k,
                    const int32_t *bounds, int ksize) {
    for (int yy = 0; yy < sh; yy++) {
        const uint8_t *row = in + (size_t)yy * sw * 4;
        uint8_t *orow = out + (size_t)yy * w * 4;
        for (int xx = 0; xx < w; xx++) {
            int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
            const int32_t *kk = k + (size_t)xx * ksize;
            int32_t s0, s1, s2, s3;
            s0 = s1 = s2 = s3 = 1 << (ORC_PRECISION_BITS - 1);
            for (int x = 0; x < xmax; x++) {
                const uint8_t *p = row + (size_t)(x + xmin) * 4;
                s0 += p[0] * kk[x];
                s1 += p[1] * kk[x];
                s2 += p[2] * kk[x];
                s3 += p[3] * kk[x];
            }
            orow[4 * xx + 0] = clip8(s0);
            orow[4 * xx + 1] = clip8(s1);
            orow[4 * xx + 2] = clip8(s2);
            orow[4 * xx + 3] = clip8(s3);
        }
    }
}

/* in: sh x w, out: h x w */
void orc_resample_v(const uint8_t *in, int w, int sh, uint8_t *out, int h, const int32_t *k,
                    const int32_t *bounds, int ksize) {
    (void)sh;
    for (int yy = 0; yy < h; yy++) {
        int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
        const int32_t *kk = k + (size_t)yy * ksize;
        uint8_t *orow = out + (size_t)yy * w * 4;
        for (int xx = 0; xx < w; xx++) {
            int32_t s0, s1, s2, s3;
            s0 = s1 = s2 = s3 = 1 << (ORC_PRECISION_BITS - 1);
            for (int y = 0; y < ymax; y++) {
                const uint8_t *p = in + ((size_t)(y + ymin) * w + xx) * 4;
                s0 += p[0] * kk[y];
                s1 += p[1] * kk[y];
                s2 += p[2] * kk[y];
                s3 += p[3] * kk[y];
            }
            orow[4 * xx + 0] = clip8(s0);
            orow[4 * xx + 1] = clip8(s1);
            orow[4 * xx + 2] = clip8(s2);
            orow[4 * xx + 3] = clip8(s3);
        }
    }
}

/* One core.resize call on an RGBa image: H pass if the width changes, then V
 * pass if the height changes, each writing a rounded uint8 image.  Returns a
 * malloc'd buffer (h x w). */
static uint8_t *core_resize(const uint8_t *in, int sw, int sh, int w, int h) {
    const uint8_t *cur = in;
    uint8_t *tmp = NULL;
    if (w != sw) {
        int ks = orc_ksize(sw, w);
        int32_t *k = (int32_t *)malloc(sizeof(int32_t) * (size_t)ks * w);
        int32_t *b = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)w);
        orc_coeffs(sw, w, k, b);
        tmp = (uint8_t *)malloc((size_t)sh * w * 4);
        orc_resample_h(cur, sw, sh, tmp, w, k, b, ks);
        free(k);
        free(b);
        cur = tmp;
    }
    uint8_t *out;
    if (h != sh) {
        int ks = orc_ksize(sh, h);
        int32_t *k = (int32_t *)malloc(sizeof(int32_t) * (size_t)ks * h);
        int32_t *b = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)h);
        orc_coeffs(sh, h, k, b);
        out = (uint8_t *)malloc((size_t)h * w * 4);
        orc_resample_v(cur, w, sh, out, h, k, b, ks);
        free(k);
        free(b);
        free(tmp);
    } else if (tmp) {
        out = tmp;
    } else {
        out = (uint8_t *)malloc((size_t)h * w * 4);
        memcpy(out, in, (size_t)h * w * 4);
    }
    return out;
}

/* Image.resize((w,h), LANCZOS) on an RGBA image (compositor.py:20 ->
 * PIL Image.py:2328-2438).  vertical_first_rule != 0 enables the Pillow 12.x
 * tall-image branch (Image.py:2431-2435). */
int orc_resize_rgba_lanczos(const uint8_t *src, int sw, int sh, uint8_t *dst, int w, int h,
                            int vertical_first_rule) {
    if (w < 1 || h < 1 || sw < 1 || sh < 1) return -1;
    if (w == sw && h == sh) { /* identity: copy BEFORE any premultiply */
        memcpy(dst, src, (size_t)sw * sh * 4);
        return 0;
    }
    uint8_t *pm = (uint8_t *)malloc((size_t)sw * sh * 4);
    orc_premultiply(src, pm, (size_t)sw * sh);
    uint8_t *res;
    if (vertical_first_rule && sh > sw * 100 && h < sh) {
        uint8_t *mid = core_resize(pm, sw, sh, sw, h);
        res = core_resize(mid, sw, h, w, h);
        free(mid);
    } else {
        res = core_resize(pm, sw, sh, w, h);
    }
    orc_unpremultiply(res, dst, (size_t)w * h);
    free(res);
    free(pm);
    return 0;
}

/* ---- AlphaComposite.c -------------------------------------------------- */

static inline uint32_t shiftfordiv255(uint32_t a) { return ((a >> 8) + a) >> 8; }

static inline void over_px(const uint8_t *dst, const uint8_t *src, uint8_t *out) {
    if (src[3] == 0) {
        out[0] = dst[0]; out[1] = dst[1]; out[2] = dst[2]; out[3] = dst[3];
        return;
    }
    uint32_t blend = dst[3] * (255u - src[3]);
    uint32_t outa255 = src[3] * 255u + blend;
    uint32_t coef1 = src[3] * 255u * 255u * (1u << 7) / outa255;
    uint32_t coef2 = 255u * (1u << 7) - coef1;
    for (int c = 0; c < 3; c++) {
        uint32_t t = src[c] * coef1 + dst[c] * coef2;
        out[c] = (uint8_t)(shiftfordiv255(t + (0x80u << 7)) >> 7);
    }
    out[3] = (uint8_t)shiftfordiv255(outa255 + 0x80u);
}

/* Image.alpha_composite(im, dest=(x,y)) in place (compositor.py:21 ->
 * PIL Image.py:1933-1987): crop (zero padded), core over, paste clipped --
 * only the intersection with the canvas changes. */
void orc_alpha_over_inplace(uint8_t *canvas, int W, int H, const uint8_t *ov, int w, int h, int x,
                            int y) {
    for (int j = 0; j < h; j++) {
        int cy = y + j;
        if (cy < 0 || cy >= H) continue;
        for (int i = 0; i < w; i++) {
            int cx = x + i;
            if (cx < 0 || cx >= W) continue;
            uint8_t *d = canvas + ((size_t)cy * W + cx) * 4;
            uint8_t o[4];
            over_px(d, ov + ((size_t)j * w + i) * 4, o);
            d[0] = o[0]; d[1] = o[1]; d[2] = o[2]; d[3] = o[3];
        }
    }
}

/* compositor.py:6-22 after the host-side id/box coercion: copy bg, then per
 * placement in list order resize to (max(1,x2-x1), max(1,y2-y1)) and over at
 * (x1,y1). */
int orc_composite(const uint8_t *bg, int W, int H, uint8_t *out, int n, const orc_placement *p,
                  int vertical_first_rule) {
    memcpy(out, bg, (size_t)W * H * 4);
    for (int i = 0; i < n; i++) {
        int w = p[i].x2 - p[i].x1;
        int h = p[i].y2 - p[i].y1;
        if (w < 1) w = 1;
        if (h < 1) h = 1;
        uint8_t *r = (uint8_t *)malloc((size_t)w * h * 4);
        if (!r) return -1;
        orc_resize_rgba_lanczos(p[i].src, p[i].sw, p[i].sh, r, w, h, vertical_first_rule);
        orc_alpha_over_inplace(out, W, H, r, w, h, p[i].x1, p[i].y1);
        free(r);
    }
    return 0;
}

/* ---- background_resizing.py -------------------------------------------- */

/* _median_color_nontransparent (background_resizing.py:11-22) restricted to the
 * rectangle [x0,x1) x [y0,y1) (the whole image for fill_solid, an 8 px edge
 * strip for _edge_strip_median_colors :36-55).  np.median over the alpha>0
 * pixels (all pixels if none), float64 mean of the two middle values, int()
 * truncation == (v[(N-1)/2] + v[N/2]) / 2 in integers. */
void orc_masked_median_rgb(const uint8_t *img, int W, int H, int x0, int y0, int x1, int y1,
                           int32_t out[3]) {
    (void)H;
    uint64_t hist[3][256];
    uint64_t n = 0;
    memset(hist, 0, sizeof hist);
    for (int pass = 0; pass < 2 && n == 0; pass++) {
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                const uint8_t *p = img + ((size_t)y * W + x) * 4;
                if (pass == 1 || p[3] > 0) {
                    hist[0][p[0]]++; hist[1][p[1]]++; hist[2][p[2]]++;
                    n++;
                }
            }
    }
    for (int c = 0; c < 3; c++) {
        if (n == 0) { out[c] = 0; continue; }
        uint64_t lo_rank = (n - 1) / 2, hi_rank = n / 2, acc = 0;
        int lo = -1, hi = -1;
        for (int v = 0; v < 256; v++) {
            acc += hist[c][v];
            if (lo < 0 && acc > lo_rank) lo = v;
            if (hi < 0 && acc > hi_rank) { hi = v; break; }
        }
        out[c] = (lo + hi) / 2;
    }
}

/* Image.new("RGBA", size, color + (255,)) (background_resizing.py:32) */
void orc_fill_rgba(uint8_t *dst, int W, int H, int r, int g, int b, int a) {
    for (size_t i = 0; i < (size_t)W * H; i++) {
        dst[4 * i + 0] = (uint8_t)r; dst[4 * i + 1] = (uint8_t)g;
        dst[4 * i + 2] = (uint8_t)b; dst[4 * i + 3] = (uint8_t)a;
    }
}

/* fill_gradient body (background_resizing.py:74-97): t = i / max(1, n-1) in
 * double, then float32(1-t)*c1 + float32(t)*c2 with separate float32 multiply
 * and add (NumPy weak-scalar promotion keeps the array dtype), astype(uint8)
 * truncation, alpha 255.  Compile with -ffp-contract=off. */
void orc_fill_gradient(uint8_t *dst, int W, int H, int horizontal, const int32_t c1[3],
                       const int32_t c2[3]) {
    int n = horizontal ? W : H;
    int den = n - 1 < 1 ? 1 : n - 1;
    for (int i = 0; i < n; i++) {
        double t = (double)i / (double)den;
        volatile float a = (float)(1.0 - t);
        volatile float b = (float)t;
        uint8_t rgb[3];
        for (int c = 0; c < 3; c++) {
            volatile float m1 = a * (float)c1[c];
            volatile float m2 = b * (float)c2[c];
            volatile float s = m1 + m2;
            rgb[c] = (uint8_t)(int)s;
        }
        if (horizontal) {
            for (int y = 0; y < H; y++) {
                uint8_t *p = dst + ((size_t)y * W + i) * 4;
                p[0] = rgb[0]; p[1] = rgb[1]; p[2] = rgb[2]; p[3] = 255;
            }
        } else {
            for (int x = 0; x < W; x++) {
                uint8_t *p = dst + ((size_t)i * W + x) * 4;
                p[0] = rgb[0]; p[1] = rgb[1]; p[2] = rgb[2]; p[3] = 255;
            }
        }
    }
}
