"""B200-native drop-in for the deterministic compositor hot path of
FelixMul/image_transformation.

Public surface (same names and semantics as the reference modules):

* ``image_transformation_b200.compositor``          -> /root/reference/compositor.py
* ``image_transformation_b200.background_resizing`` -> /root/reference/background_resizing.py
* ``image_transformation_b200.batch``               -> device-resident batched API (many
  independent canvases per launch, sharded by canvas across GPUs)
* ``image_transformation_b200.sheets``              -> the same two raster operations where the reference uses
  them outside ``composite()``: contact sheet and candidates grid (macro_placement_test.py:162-242, 1332-1345),
  RGB downscale for VLM uploads (api_client.py:97-112)

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/b200comp.h`` (``_lib/libb200comp.so``); there is no CPU fallback.
"""
__version__ = "0.2.0"

import os as _os


def _raise_pillow_block_size() -> None:
    """Pillow stores an image in blocks of PILLOW_BLOCK_SIZE (16 MB by default) and can hand out an image's memory
    without a copy only when it lives in ONE block.  A 4K RGBA canvas is 33 MB: with the default every composite()
    would first copy its background into a single block (about 10 ms).  Images allocated after this call (the
    caller's Image.open(...).convert("RGBA") included) use blocks of up to 512 MB, i.e. one block per image; small
    images are unaffected (Pillow allocates what an image needs, not a whole block).  B200COMP_PILLOW_BLOCK=0 leaves
    Pillow's setting alone."""
    if _os.environ.get("B200COMP_PILLOW_BLOCK", "1") == "0":
        return
    try:
        from PIL import Image

        if Image.core.get_block_size() < (512 << 20):
            Image.core.set_block_size(512 << 20)
    except Exception:  # pragma: no cover - an older Pillow without the allocator knobs
        pass


_raise_pillow_block_size()

from . import _native  # noqa: F401  (fails loudly if the library is missing and cannot be built)
