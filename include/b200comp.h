/*
 * b200comp.h -- C ABI of the B200-native compositor hot path (libb200comp.so).
 *
 * Drop-in boundary for the deterministic compositor of FelixMul/image_transformation:
 * every entry point names the reference interface it replaces (file:line under
 * /root/reference, or the Pillow / NumPy routine that interface delegates to).
 * Plain pointers and sizes only; no torch / PIL types.  All pixel buffers are
 * uint8 RGBA, interleaved, row-major, 4 bytes per pixel (PIL mode "RGBA");
 * pitches are in BYTES.  "device" pointers are CUDA device pointers on the
 * calling thread's current device; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).
 *
 * Return value: 0 on success, negative b200comp_status on failure;
 * b200comp_last_error() then returns a thread-local message.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with
 * B200COMP_ECUDA.
 */
#ifndef B200COMP_H
#define B200COMP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200COMP_ABI_VERSION 1

enum b200comp_status {
    B200COMP_OK = 0,
    B200COMP_EINVAL = -1, /* bad argument (NULL pointer, non-positive size, ...) */
    B200COMP_ECUDA = -2,  /* CUDA runtime error / no device */
    B200COMP_ENOMEM = -3, /* host or device allocation failed */
    B200COMP_EINTERNAL = -4
};

/* Placement flag: run the vertical pass before the horizontal one.  The host
 * mirror sets it when Pillow >= 12 would (src_h > 100*src_w and h < src_h,
 * PIL Image.py:2431-2435, reached from compositor.py:20). */
#define B200COMP_VERTICAL_FIRST 1
/* Placement flag for the host-buffer entry points: `src` is DEVICE memory (a cutout uploaded once with
 * b200comp_device_upload and reused over many calls -- the refine loop of macro_placement_test.py:1493-1513,
 * 1679-1699 re-composites the same bundle every iteration). */
#define B200COMP_SRC_DEVICE 2
/* Placement flag: the resampled cutout REPLACES the canvas pixels of its box instead of being composited onto them
 * (`obj.resize((w, h), LANCZOS)` without the alpha_composite that follows it in compositor.py:20-21; Convert.c
 * rgba2rgbA keeps the colours of pixels whose alpha resampled to 0).  Only for placements the fused kernel
 * resamples (size != cutout size, scale down to 2.66x, no vertical-first order); b200comp_plan_create rejects others. */
#define B200COMP_REPLACE 4

int b200comp_abi_version(void);
const char *b200comp_last_error(void);
/* Number of visible CUDA devices (0 if none); never fails. */
int b200comp_device_count(void);

/* ------------------------------------------------------------------------
 * Coefficient tables (HOST, double precision + libm sin).
 * Replaces Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc with the
 * LANCZOS filter (support 3), reached from `obj.resize((w, h), Image.LANCZOS)`
 * at compositor.py:20.
 *   ksize  = 2*ceil(3*max(1, in/out)) + 1
 *   k      : out_size * ksize int32, 22-bit fixed point, zero padded past xmax
 *   bounds : out_size * 2 int32 (xmin, xmax)
 * ---------------------------------------------------------------------- */
int b200comp_ksize(int in_size, int out_size);
int b200comp_build_coeffs(int in_size, int out_size, int32_t *k_host, int32_t *bounds_host, int *ksize);

/* ------------------------------------------------------------------------
 * Stand-alone separable resampler (device buffers).
 * Replaces `obj.resize((w, h), Image.LANCZOS)` for an RGBA cutout
 * (compositor.py:20 -> PIL Image.py:2328-2438): identity copy, else
 * premultiply -> H pass -> uint8 -> V pass -> un-premultiply.
 * Builds and uploads the coefficient tables itself (stream ordered).
 * ---------------------------------------------------------------------- */
int b200comp_resize_rgba_lanczos(const uint8_t *src, int sw, int sh, size_t src_pitch, uint8_t *dst, int w,
                                 int h, size_t dst_pitch, int flags, void *stream);

/* Same, with caller-provided DEVICE coefficient tables (from b200comp_build_coeffs,
 * copied to the device by the caller).  A pass whose size does not change must be
 * given ks = 0 and NULL tables.  `scratch` is a device buffer of at least
 * max(sh*w, h*sw)*4 bytes (the uint8 intermediate). */
int b200comp_resample_rgba(const uint8_t *src, int sw, int sh, size_t src_pitch, uint8_t *dst, int w, int h,
                           size_t dst_pitch, const int32_t *kx, const int32_t *bx, int ksx, const int32_t *ky,
                           const int32_t *by, int ksy, uint8_t *scratch, int flags, void *stream);

/* ------------------------------------------------------------------------
 * In-place alpha-over of an RGBA overlay at dest=(x, y) (device buffers).
 * Replaces `canvas.alpha_composite(resized, dest=(x1, y1))` (compositor.py:21 ->
 * PIL Image.py:1933-1987 -> AlphaComposite.c): only the intersection with the
 * canvas changes; negative / overhanging dest clips.
 * ---------------------------------------------------------------------- */
int b200comp_alpha_over(uint8_t *canvas, int W, int H, size_t pitch, const uint8_t *src, int w, int h,
                        size_t src_pitch, int x, int y, void *stream);

/* ------------------------------------------------------------------------
 * Fused batched compositor (device buffers): for every canvas, copy the
 * background (or synthesise the solid colour) and resample + alpha-over its
 * placements in list order, one pass over each output tile.
 * Replaces the whole body of composite() (compositor.py:11-22) for a batch of
 * independent canvases.  Host-side coercions of compositor.py:13-18 (id lookup,
 * int() truncation, max(1, .)) are done by the caller: (x, y) = (x1, y1),
 * (w, h) = (max(1, x2-x1), max(1, y2-y1)).
 * ---------------------------------------------------------------------- */
typedef struct b200comp_placement {
    const uint8_t *src; /* cutout pixels */
    int64_t src_pitch;
    int32_t sw, sh;     /* cutout size */
    int32_t x, y, w, h; /* destination top-left and resampled size (w, h >= 1) */
    int32_t flags;      /* B200COMP_VERTICAL_FIRST */
    int32_t reserved;
} b200comp_placement;

typedef struct b200comp_canvas {
    uint8_t *out; /* W x H output */
    int64_t out_pitch;
    const uint8_t *bg; /* W x H background, or NULL: use solid_rgba */
    int64_t bg_pitch;
    uint32_t solid_rgba; /* R | G<<8 | B<<16 | A<<24 (byte order of an RGBA pixel) */
    int32_t W, H;
    int32_t first_placement; /* index into the placement array */
    int32_t n_placements;    /* z-order = array order, later on top */
    int32_t reserved;
} b200comp_canvas;

typedef struct b200comp_plan b200comp_plan;

/* Resolve a batch: de-duplicate and build the coefficient tables on
 * `n_host_threads` host threads (<= 0: hardware concurrency), upload them and
 * the descriptors (stream ordered).  Pixel pointers in the descriptors are
 * DEVICE pointers and must stay valid until the plan is destroyed. */
int b200comp_plan_create(const b200comp_canvas *canvases, int n_canvases, const b200comp_placement *placements,
                         int n_placements, int n_host_threads, void *stream, b200comp_plan **plan);
/* Launch the batch (asynchronous on `stream`).  Re-runnable.  Equivalent to b200comp_plan_prepare
 * followed by b200comp_plan_run_canvases over every canvas.  Stream semantics are the caller's stream's: the work
 * starts after everything queued on `stream` before the call and is complete for everything queued after it.  A large
 * run is cut into waves whose binning runs on plan-owned side streams under the previous wave's tile kernel; the side
 * streams fork from and join back into `stream` with events before the call returns (B200COMP_WAVES=1 switches the
 * split off).  Runs of one plan share its buffers and must be ordered (same stream, or synchronised). */
int b200comp_plan_run(b200comp_plan *plan, void *stream);
/* The two stages separately, so a caller can pipeline copies with compute: `prepare` builds what the
 * tile kernel reads besides the caller's buffers (premultiplied planar cutouts, pre-resampled overlays
 * of extreme scales); `run_canvases` launches the fused kernel for canvases [first, first + count). */
int b200comp_plan_prepare(b200comp_plan *plan, void *stream);
int b200comp_plan_run_canvases(b200comp_plan *plan, int first, int count, void *stream);
/* Frees the plan's device memory stream-ordered on the stream it was created on; the caller
 * must have synchronised every stream the plan ran on. */
int b200comp_plan_destroy(b200comp_plan *plan);

enum b200comp_plan_info_key {
    B200COMP_INFO_ALGORITHMIC_BYTES = 0, /* sum over canvases of bg read + out write + each placed cutout once */
    B200COMP_INFO_LAUNCHES_PER_RUN = 1,  /* kernels launched by one b200comp_plan_run */
    B200COMP_INFO_FUSED_PLACEMENTS = 2,  /* resampled inside the fused tile kernel */
    B200COMP_INFO_IDENTITY_PLACEMENTS = 3,
    B200COMP_INFO_PRERESAMPLED_PLACEMENTS = 4, /* extreme scales: generic two-pass kernels first */
    B200COMP_INFO_COEFF_BYTES = 5,
    B200COMP_INFO_SMEM_BYTES = 6,
    B200COMP_INFO_TILES = 7,
    B200COMP_INFO_COUNT = 8
};
int b200comp_plan_info(const b200comp_plan *plan, int64_t *info /* [B200COMP_INFO_COUNT] */);
/* Synchronise `stream` and read the kernel's status word: fails with B200COMP_EINTERNAL if a
 * tile needed more shared memory than the plan sized (never silently wrong pixels). */
int b200comp_plan_check(b200comp_plan *plan, void *stream);
/* Synchronise `stream` and return the number of command records the binning pass of the last
 * b200comp_plan_run[_canvases] wrote: one per tile something is drawn on, one per (tile, placement)
 * step that survived the transparency and occlusion tests, one END per persistent CTA. */
int b200comp_plan_last_records(b200comp_plan *plan, void *stream, int64_t *records);
/* Measurement aid (bench.py): with profiling enabled every b200comp_plan_run[_canvases] brackets the
 * prepare kernel, the three binning kernels and the tile kernel with CUDA events on the launching
 * stream.  b200comp_plan_profile_read synchronises the stream of the last run and returns the summed
 * durations in milliseconds: ms[0] prepare, ms[1] binning, ms[2] tile kernel; *runs = runs measured;
 * the counters are reset. */
int b200comp_plan_profile(b200comp_plan *plan, int enable);
int b200comp_plan_profile_read(b200comp_plan *plan, double ms[3], int *runs);

/* create + run + destroy */
int b200comp_composite_batch(const b200comp_canvas *canvases, int n_canvases, const b200comp_placement *placements,
                             int n_placements, void *stream);

/* ------------------------------------------------------------------------
 * Host-buffer entry points (what a ctypes / cffi binding of the reference
 * calls): same descriptors, but every pixel pointer is a HOST pointer
 * (pinned memory overlaps copies with compute; pageable works).  Copies
 * cutouts (de-duplicated by pointer) and backgrounds to the device, runs the
 * fused kernel in chunks of `chunk_canvases` canvases (<= 0: about 32 MB of canvas per chunk) over three
 * streams (`n_streams` is kept for ABI stability) and copies the canvases back.
 * Synchronous.  b200comp_composite_host is the single-canvas form behind the
 * drop-in composite() (compositor.py:6).
 * ---------------------------------------------------------------------- */
int b200comp_composite_batch_host(const b200comp_canvas *canvases, int n_canvases,
                                  const b200comp_placement *placements, int n_placements, int n_host_threads,
                                  int chunk_canvases, int n_streams);
int b200comp_composite_host(const uint8_t *bg, int W, int H, size_t bg_pitch, uint8_t *out, size_t out_pitch,
                            const b200comp_placement *placements, int n_placements);
/* Same call with two short cuts for the callers around composite() (macro_placement_test.py:1427 -> :1510-1511):
 * bg == NULL composites onto a canvas of the solid colour `solid_rgba` (what fill_solid returned: no W*H*4 upload),
 * and placements flagged B200COMP_SRC_DEVICE read cutouts that already live on the device.  Both forms run on a
 * per-thread cached context (one stream, pinned bounce buffers, device staging): no helper thread, no device-wide
 * synchronisation, safe to call from several threads at once. */
int b200comp_composite_host_ex(const uint8_t *bg, uint32_t solid_rgba, int W, int H, size_t bg_pitch, uint8_t *out,
                               size_t out_pitch, const b200comp_placement *placements, int n_placements);
/* Device-resident copy of a decoded RGBA image (16-byte aligned pitch); free with b200comp_device_free. */
int b200comp_device_upload(const uint8_t *img, int w, int h, size_t pitch, uint8_t **dev, size_t *dev_pitch);
int b200comp_device_free(uint8_t *dev);
/* Give cached memory back to the driver: the library's own stream-ordered memory pool (created per device; the
 * process's default pool and its release threshold are never touched) and the calling thread's cached host-call
 * context (pinned bounce buffers, device staging). */
int b200comp_trim(void);
/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
int b200comp_host_alloc(void **ptr, size_t bytes);
int b200comp_host_free(void *ptr);

/* ------------------------------------------------------------------------
 * Background statistics and synthesis (device buffers).
 * ---------------------------------------------------------------------- */
/* Replaces _median_color_nontransparent (background_resizing.py:11-22) on the
 * rectangle [x0,x1) x [y0,y1): per-channel median over alpha>0 pixels (all
 * pixels if there are none), (v[(N-1)/2] + v[N/2]) / 2.  out_rgb is a HOST
 * array of 3 ints; the call synchronises `stream`.  The 8 px edge strips of
 * _edge_strip_median_colors (:36-55) are four calls with different rectangles. */
int b200comp_masked_median_rgb(const uint8_t *img, int W, int H, size_t pitch, int x0, int y0, int x1, int y1,
                               int32_t out_rgb[3], void *stream);
/* Replaces Image.new("RGBA", size, color + (255,)) (background_resizing.py:32). */
int b200comp_fill_rgba(uint8_t *dst, int W, int H, size_t pitch, uint32_t rgba, void *stream);
/* Replaces the loop of fill_gradient (background_resizing.py:74-97): float32 lerp
 * from c1 to c2 along x (horizontal != 0) or y, truncated, alpha 255. */
int b200comp_fill_gradient(uint8_t *dst, int W, int H, size_t pitch, int horizontal, const int32_t c1[3],
                           const int32_t c2[3], void *stream);
/* Host-buffer forms of the statistics: img is a decoded RGBA image in HOST memory.
 * Replace _median_color_nontransparent (:11-22) and _edge_strip_median_colors (:36-55;
 * out_edges = left, right, top, bottom as 4 x (r, g, b)). */
int b200comp_masked_median_rgb_host(const uint8_t *img, int W, int H, size_t pitch, int x0, int y0, int x1, int y1,
                                    int32_t out_rgb[3]);
int b200comp_edge_strip_medians_host(const uint8_t *img, int W, int H, size_t pitch, int strip_px,
                                     int32_t out_edges[12]);
/* Host-buffer forms behind fill_solid / fill_gradient (background_resizing.py:25,63):
 * bg is the decoded background.png (HOST), out the canvas (HOST). */
int b200comp_fill_solid_host(const uint8_t *bg, int Wb, int Hb, size_t bg_pitch, uint8_t *out, int W, int H,
                             size_t out_pitch, int32_t out_rgb[3]);
int b200comp_fill_gradient_host(const uint8_t *bg, int Wb, int Hb, size_t bg_pitch, uint8_t *out, int W, int H,
                                size_t out_pitch, int strip_px, int32_t out_edges[12], int *out_horizontal);

#ifdef __cplusplus
}
#endif
#endif /* B200COMP_H */
