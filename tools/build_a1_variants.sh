#!/bin/bash
# A/B builds of the type A1 Miller kernel for tools/gpu_variants.sh (run HERE, nvcc cross-compiles):
#   build/liba1_default.so   14 slots, 96 threads per block (the measured configuration)
#   build/liba1_s13.so       PBC_A1_SLOTS13=1: five-temporary slot programs, 13 slots, 128 threads
#   build/liba1_naf.so       PBC_A1_NAF=1: signed-digit scan of n (a third fewer chord steps)
#   build/liba1_s13naf.so    both
#   build/libcc_naf.so       PBC_CC_NAF=1: signed-digit scan of r in the type F / D / G Miller loop (-6.7 % / -6.9 % /
#                            -3.6 % multiplier work per pairing by the simulator's count); WL="f 131072" etc.
# then on the GPU box:
#   VARIANTS="a1_default a1_s13 a1_naf a1_s13naf" WL="a1 28416" bash tools/gpu_variants.sh
#   for v in a1_s13 a1_naf a1_s13naf; do PBC_B200_LIB=$PWD/build/lib$v.so python -m pytest tests/test_gpu_type_a1.py -q; done
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC"
nvcc $FLAGS -o build/liba1_default.so pbc_b200/csrc/engine.cu -lcudart
nvcc $FLAGS -DPBC_A1_SLOTS13=1 -o build/liba1_s13.so pbc_b200/csrc/engine.cu -lcudart
nvcc $FLAGS -DPBC_A1_NAF=1 -o build/liba1_naf.so pbc_b200/csrc/engine.cu -lcudart
nvcc $FLAGS -DPBC_A1_SLOTS13=1 -DPBC_A1_NAF=1 -o build/liba1_s13naf.so pbc_b200/csrc/engine.cu -lcudart
nvcc $FLAGS -DPBC_CC_NAF=1 -o build/libcc_naf.so pbc_b200/csrc/engine.cu -lcudart
ls -la build/liba1_*.so build/libcc_naf.so
