"""Device-resident batched compositor: many independent canvases per launch.

This is the throughput API behind the benchmark (BASELINE.json configs C3-C5): the cutout
pool, the backgrounds and the outputs live in HBM, one `b200comp_plan` resolves the whole
batch (coefficient tables built once on host threads) and every `run()` is a single launch of
the fused resample + alpha-over tile kernel.  Canvases are independent, so multi-GPU use is
"one process per GPU, each with its own contiguous block of canvases" (`shard_range`); there
is no collective.

PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import time
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native
from .compositor import resolve_placements


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of canvas indices owned by `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return lo, hi


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


def _stream_handle(stream: Optional[torch.cuda.Stream]) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


class CutoutPool:
    """RGBA cutouts resident in HBM: one uint8 buffer, 16-byte aligned pitches (TMA-ready),
    256-byte aligned bases."""

    def __init__(self, cutouts: Dict[int, np.ndarray], device: Union[str, torch.device, None] = None,
                 pin: bool = True):
        _native.require_gpu()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.sizes: Dict[int, Tuple[int, int]] = {}
        self._off: Dict[int, int] = {}
        self._pitch: Dict[int, int] = {}
        total = 0
        for oid, a in cutouts.items():
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4:
                raise ValueError("images do not match")  # same complaint as a non-RGBA cutout in the reference
            sh, sw = a.shape[:2]
            self.sizes[int(oid)] = (sw, sh)
            self._pitch[int(oid)] = _align(sw * 4, 16)
            self._off[int(oid)] = total
            total = _align(total + self._pitch[int(oid)] * sh, 256)
        host = torch.zeros(max(total, 16), dtype=torch.uint8, pin_memory=pin)
        hv = host.numpy()
        for oid, a in cutouts.items():
            sh, sw = a.shape[:2]
            p, o = self._pitch[int(oid)], self._off[int(oid)]
            hv[o:o + p * sh].reshape(sh, p)[:, : sw * 4] = np.ascontiguousarray(a).reshape(sh, sw * 4)
        self.h2d_bytes = int(total)
        self.buffer = host.to(self.device, non_blocking=False)

    def ptr(self, oid: int) -> int:
        return self.buffer.data_ptr() + self._off[oid]

    def pitch(self, oid: int) -> int:
        return self._pitch[oid]

    def view(self, oid: int) -> torch.Tensor:
        sw, sh = self.sizes[oid]
        p, o = self._pitch[oid], self._off[oid]
        return self.buffer[o:o + p * sh].view(sh, p)[:, : sw * 4].reshape(sh, sw, 4)


class CompositeBatch:
    """A resolved batch of canvases on one GPU.

    canvas_sizes : [(W, H)] per canvas
    placements   : per canvas, the reference's placement list ({object_id, box}); same host
                   coercions as ``compositor.composite`` (unknown ids skipped, int() truncation)
    backgrounds  : None (use ``solid``), a uint8 CUDA tensor (N, H, W, 4) or a list of (H, W, 4)
                   CUDA tensors (entries may be None -> solid)
    solid        : per-canvas (r, g, b, a) used when a canvas has no background tensor
    """

    def __init__(self, pool: CutoutPool, canvas_sizes: Sequence[Tuple[int, int]],
                 placements: Sequence[Sequence[dict]],
                 backgrounds: Union[None, torch.Tensor, Sequence[Optional[torch.Tensor]]] = None,
                 solid: Union[None, Tuple[int, int, int, int], Sequence[Tuple[int, int, int, int]]] = None,
                 out: Optional[torch.Tensor] = None, host_threads: int = 0,
                 stream: Optional[torch.cuda.Stream] = None):
        _native.require_gpu()
        n = len(canvas_sizes)
        if n < 1 or len(placements) != n:
            raise ValueError("canvas_sizes and placements must have the same non-zero length")
        self.pool = pool
        self.n = n
        self.sizes = [(int(w), int(h)) for w, h in canvas_sizes]
        self._plan = ctypes.c_void_p()
        # ---- outputs: one flat buffer, 256-byte aligned canvases, pitch = W*4 ----
        offs, total = [], 0
        for (w, h) in self.sizes:
            offs.append(total)
            total = _align(total + w * 4 * h, 256)
        if out is None:
            out = torch.empty(total, dtype=torch.uint8, device=pool.device)
        elif out.numel() < total or out.dtype != torch.uint8 or not out.is_cuda:
            raise ValueError("out must be a uint8 CUDA tensor of at least %d bytes" % total)
        self.out_buffer = out
        self._out_off = offs
        self.out_bytes = sum(w * 4 * h for w, h in self.sizes)
        # ---- descriptors ----
        if solid is None:
            solid = (0, 0, 0, 255)
        solids = [solid] * n if isinstance(solid[0], int) else list(solid)
        recs: List[tuple] = []
        cv = (_native.Canvas * n)()
        self._bg_keepalive = backgrounds
        self.bg_bytes = 0
        for i, (w, h) in enumerate(self.sizes):
            res = resolve_placements(placements[i], pool.sizes)
            first = len(recs)
            recs.extend(res)
            bg_t = None
            if isinstance(backgrounds, torch.Tensor):
                bg_t = backgrounds[i]
            elif backgrounds is not None:
                bg_t = backgrounds[i]
            bg_ptr, bg_pitch = 0, 0
            if bg_t is not None:
                if bg_t.dtype != torch.uint8 or tuple(bg_t.shape) != (h, w, 4) or not bg_t.is_cuda:
                    raise ValueError("image has wrong mode")  # not an RGBA canvas of the right size
                if bg_t.stride(2) != 1 or bg_t.stride(1) != 4:
                    raise ValueError("background rows must be contiguous RGBA")
                bg_ptr, bg_pitch = bg_t.data_ptr(), bg_t.stride(0)
                self.bg_bytes += w * 4 * h
            r, g, b, a = (int(v) & 0xFF for v in solids[i])
            cv[i] = _native.Canvas(out.data_ptr() + offs[i], w * 4, bg_ptr or None, bg_pitch,
                                   r | (g << 8) | (b << 16) | (a << 24), w, h, first, len(res), 0)
        pl = (_native.Placement * max(1, len(recs)))()
        self.src_bytes = 0
        for j, (oid, x, y, w, h, flags) in enumerate(recs):
            sw, sh = pool.sizes[oid]
            pl[j] = _native.Placement(pool.ptr(oid), pool.pitch(oid), sw, sh, x, y, w, h, flags, 0)
            self.src_bytes += sw * sh * 4
        self.n_placements = len(recs)
        self._cv, self._pl, self._host_threads = cv, pl, int(host_threads)  # kept for recreate()
        self.plan_create_s = 0.0
        self._create(stream)

    def _create(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        t0 = time.perf_counter()
        with torch.cuda.device(self.pool.device):
            rc = _native.lib().b200comp_plan_create(self._cv, self.n, self._pl, self.n_placements, self._host_threads,
                                                    _stream_handle(stream), ctypes.byref(self._plan))
        _native.check(rc, "CompositeBatch")
        self.plan_create_s = time.perf_counter() - t0  # the C call alone: tables, tensor maps, uploads
        info = (ctypes.c_int64 * len(_native.INFO_KEYS))()
        _native.check(_native.lib().b200comp_plan_info(self._plan, info), "plan_info")
        self.info = dict(zip(_native.INFO_KEYS, (int(v) for v in info)))

    def recreate(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Throw the plan away and resolve the same descriptors again (what a fresh layout costs: coefficient
        tables, tensor maps, uploads) -- bench.py's fresh-plan figure."""
        if self._plan:
            with torch.cuda.device(self.pool.device):
                _native.lib().b200comp_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()
        self._create(stream)

    # ------------------------------------------------------------------ execution
    def run(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """One launch of the fused tile kernel (asynchronous on the current torch stream)."""
        with torch.cuda.device(self.pool.device):
            _native.check(_native.lib().b200comp_plan_run(self._plan, _stream_handle(stream)), "CompositeBatch.run")

    def run_canvases(self, first: int, count: int, stream: Optional[torch.cuda.Stream] = None, prepare: bool = True) -> None:
        """Composite canvases [first, first + count) only (b200comp_plan_prepare + b200comp_plan_run_canvases): what a
        caller pipelining copies with compute does chunk by chunk.  `prepare=False` when the cutouts were prepared by an
        earlier call on the same stream and have not changed."""
        with torch.cuda.device(self.pool.device):
            if prepare:
                _native.check(_native.lib().b200comp_plan_prepare(self._plan, _stream_handle(stream)), "CompositeBatch.prepare")
            _native.check(_native.lib().b200comp_plan_run_canvases(self._plan, int(first), int(count), _stream_handle(stream)),
                          "CompositeBatch.run_canvases")

    def check(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Synchronise and verify the kernel's status word."""
        with torch.cuda.device(self.pool.device):
            _native.check(_native.lib().b200comp_plan_check(self._plan, _stream_handle(stream)), "CompositeBatch.check")

    def last_records(self, stream: Optional[torch.cuda.Stream] = None) -> int:
        """Command records written by the binning pass of the last run (tiles + surviving steps + one END per CTA)."""
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.pool.device):
            _native.check(_native.lib().b200comp_plan_last_records(self._plan, _stream_handle(stream), ctypes.byref(n)),
                          "CompositeBatch.last_records")
        return int(n.value)

    def profile(self, enable: bool = True) -> None:
        """Bracket the phases of every following run() with CUDA events on the launching stream."""
        _native.check(_native.lib().b200comp_plan_profile(self._plan, int(bool(enable))), "CompositeBatch.profile")

    def profile_read(self) -> dict:
        """Summed phase durations (ms) of the runs since profile(True); synchronises the stream."""
        ms = (ctypes.c_double * 3)()
        runs = ctypes.c_int(0)
        with torch.cuda.device(self.pool.device):
            _native.check(_native.lib().b200comp_plan_profile_read(self._plan, ms, ctypes.byref(runs)), "CompositeBatch.profile_read")
        return {"runs": int(runs.value), "prepare_ms": float(ms[0]), "binning_ms": float(ms[1]), "tile_kernel_ms": float(ms[2])}

    def output(self, i: int) -> torch.Tensor:
        w, h = self.sizes[i]
        o = self._out_off[i]
        return self.out_buffer[o:o + w * 4 * h].view(h, w, 4)

    def outputs(self) -> List[torch.Tensor]:
        return [self.output(i) for i in range(self.n)]

    @property
    def algorithmic_bytes(self) -> int:
        """bg read + out write + every placed cutout read once."""
        return self.info["algorithmic_bytes"]

    def close(self) -> None:
        if self._plan:
            torch.cuda.synchronize(self.pool.device)
            _native.lib().b200comp_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------
# thin device-level wrappers over the single-op C ABI (used by the stage parity tests)
# ---------------------------------------------------------------------------------------
def _chk_img(t: torch.Tensor) -> None:
    if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 4 and t.stride(2) == 1 and t.stride(1) == 4):
        raise ValueError("expected a uint8 CUDA tensor (H, W, 4) with contiguous rows")


def resize_rgba_lanczos(src: torch.Tensor, size: Tuple[int, int], vertical_first: bool = False) -> torch.Tensor:
    """`Image.resize(size, LANCZOS)` of an RGBA tensor on the device; size = (w, h)."""
    _chk_img(src)
    w, h = int(size[0]), int(size[1])
    dst = torch.empty((h, w, 4), dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        rc = _native.lib().b200comp_resize_rgba_lanczos(src.data_ptr(), src.shape[1], src.shape[0], src.stride(0),
                                                        dst.data_ptr(), w, h, dst.stride(0),
                                                        _native.VERTICAL_FIRST if vertical_first else 0,
                                                        _stream_handle(None))
    _native.check(rc, "resize_rgba_lanczos")
    return dst


def alpha_over_(canvas: torch.Tensor, overlay: torch.Tensor, dest: Tuple[int, int]) -> torch.Tensor:
    """In-place `canvas.alpha_composite(overlay, dest)` on the device."""
    _chk_img(canvas)
    _chk_img(overlay)
    with torch.cuda.device(canvas.device):
        rc = _native.lib().b200comp_alpha_over(canvas.data_ptr(), canvas.shape[1], canvas.shape[0], canvas.stride(0),
                                               overlay.data_ptr(), overlay.shape[1], overlay.shape[0],
                                               overlay.stride(0), int(dest[0]), int(dest[1]), _stream_handle(None))
    _native.check(rc, "alpha_over_")
    return canvas


def masked_median_rgb(img: torch.Tensor, rect: Optional[Tuple[int, int, int, int]] = None) -> Tuple[int, int, int]:
    _chk_img(img)
    H, W = img.shape[:2]
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, W, H)
    out = (ctypes.c_int32 * 3)()
    with torch.cuda.device(img.device):
        rc = _native.lib().b200comp_masked_median_rgb(img.data_ptr(), W, H, img.stride(0), x0, y0, x1, y1, out,
                                                      _stream_handle(None))
    _native.check(rc, "masked_median_rgb")
    return int(out[0]), int(out[1]), int(out[2])


def fill_rgba_(dst: torch.Tensor, rgba: Tuple[int, int, int, int]) -> torch.Tensor:
    _chk_img(dst)
    r, g, b, a = (int(v) & 0xFF for v in rgba)
    with torch.cuda.device(dst.device):
        rc = _native.lib().b200comp_fill_rgba(dst.data_ptr(), dst.shape[1], dst.shape[0], dst.stride(0),
                                              r | (g << 8) | (b << 16) | (a << 24), _stream_handle(None))
    _native.check(rc, "fill_rgba_")
    return dst


def fill_gradient_(dst: torch.Tensor, horizontal: bool, c1: Sequence[int], c2: Sequence[int]) -> torch.Tensor:
    _chk_img(dst)
    a1 = (ctypes.c_int32 * 3)(*[int(v) for v in c1])
    a2 = (ctypes.c_int32 * 3)(*[int(v) for v in c2])
    with torch.cuda.device(dst.device):
        rc = _native.lib().b200comp_fill_gradient(dst.data_ptr(), dst.shape[1], dst.shape[0], dst.stride(0),
                                                  int(bool(horizontal)), a1, a2, _stream_handle(None))
    _native.check(rc, "fill_gradient_")
    return dst
