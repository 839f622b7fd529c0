#!/usr/bin/env bash
# Compile-check the build-time variants of the tangent kernel so they do not rot (CPU only, nvcc cross-compiles sm_100a):
#   c4    column-split pair owners, 168 registers (kernel_mat2c.cuh)
#   w22   warp-specialised persistent kernel, 2 + 2 teams, 168 registers (kernel_mat2w.cuh)
#   ko    k_mat2 with the phase knock-out predicates (tools/ko_sweep.py)
#   mat3  warp-specialised persistent kernel with 255-register consumers (tools/variants/kernel_mat3.cuh), wait / phase counters on
# usage: bash tools/build_variants.sh      (objects under finiteelementcontainers.jl_b200/build/<variant>/, libraries in lib/)
set -e
cd "$(dirname "$0")/.."
FECB200_DEFINES="-DFEC_MAT2C=1 -DFEC_MAT2C_MINB=4" FECB200_VARIANT=c4 python finiteelementcontainers.jl_b200/build.py
FECB200_DEFINES="-DFEC_MAT2W=1 -DFEC_MAT2W_PT=2 -DFEC_MAT2W_CT=2 -DFEC_MAT2W_REG=168" FECB200_VARIANT=w22 python finiteelementcontainers.jl_b200/build.py
FECB200_DEFINES="-DFEC_MAT2_KO" FECB200_VARIANT=ko python finiteelementcontainers.jl_b200/build.py
FECB200_DEFINES="-DFEC_MAT3=1 -DFEC_MAT3_PROF" FECB200_VARIANT=mat3 python finiteelementcontainers.jl_b200/build.py
echo "variants built: lib/libfecb200_{c4,w22,ko,mat3}.so  (select one with FECB200_LIB=...)"
