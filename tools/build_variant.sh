#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags...]  ->  image_transformation_b200/_lib/variants/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=image_transformation_b200/_lib/variants; mkdir -p $out/obj_$name
src=image_transformation_b200/csrc
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -fvisibility=hidden -c $src/coeffs.cpp -o $out/obj_$name/coeffs.o
for f in b200comp host_api; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -Xptxas -v "$@" -c $src/$f.cu -o $out/obj_$name/$f.o 2>&1 | grep -A2 "composite_slab" | grep -E "spill|Used" || true
done
nvcc -shared -o $out/$name.so $out/obj_$name/*.o -cudart static -lpthread 2>/dev/null
echo built $out/$name.so
