#!/bin/bash
# A/B builds for the round-2 kernels (run HERE; the .so files travel with the snapshot):
#   build/libr2_default.so    the shipped configuration
#   build/libr2_alock.so      PBC_A_LOCKSTEP=1: block-wide barrier per Miller iteration in k_a_miller9
#   build/libr2_fs96.so       PBC_FS_MILLER_BLOCK=96: three 96-thread blocks per SM in k_f_miller_s
#   build/libr2_a14.so        PBC_A_SLOTS9=0: the 14-slot type A kernel of round 1
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $FLAGS $2 -o build/libr2_$1.so pbc_b200/csrc/engine.cu -lcudart & }
build default ""
build alock "-DPBC_A_LOCKSTEP=1"
build fs96 "-DPBC_FS_MILLER_BLOCK=96"
build a14 "-DPBC_A_SLOTS9=0"
wait
ls -la build/libr2_*.so
