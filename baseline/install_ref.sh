#!/bin/bash
# Copies the UNMODIFIED reference (FelixMul/image_transformation, /root/reference in the build container) to
# baseline/_ref/ -- git-ignored, but shipped to the GPU box with the gpurun snapshot -- so that
#   * tests/test_gpu_dropin.py can run the reference's own tests/test_compositor.py against the drop-in modules, and
#   * bench.py can time the reference's own composite() (Python over Pillow) beside the GPU path.
# The reference is a directory of Python modules, not an installable package (no setup.py / pyproject.toml:
# `pip install --target baseline/_ref /root/reference` has nothing to build), hence a plain copy.
set -e
cd "$(dirname "$0")/.."
SRC=${1:-/root/reference}
[ -d "$SRC" ] || { echo "no reference at $SRC"; exit 1; }
rm -rf baseline/_ref
mkdir -p baseline/_ref
(cd "$SRC" && tar --exclude=.git --exclude='__pycache__' -cf - .) | (cd baseline/_ref && tar -xf -)
echo "reference copied to baseline/_ref ($(find baseline/_ref -type f | wc -l) files)"
