#!/bin/bash
# usage (on the GPU box): tools/gpu_phase.sh <variant> [canvases]  -- phase-cycle profile of a -DB200COMP_PROFILE=1 variant
V=${1:-prof}; N=${2:-64}
B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$V.so timeout 300 python tools/phase_profile.py $N
