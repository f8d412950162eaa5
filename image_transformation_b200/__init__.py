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
__version__ = "0.1.0"

from . import _native  # noqa: F401  (fails loudly if the library is missing and cannot be built)
