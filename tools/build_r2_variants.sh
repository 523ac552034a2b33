#!/bin/bash
# A/B builds for the round-2 kernels (run HERE; the .so files travel with the snapshot):
#   build/libr2_default.so    the shipped configuration
#   build/libr2_fsnolock.so   PBC_FS_LOCKSTEP=0: type F slot kernels without the block-wide barriers
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $FLAGS $2 -o build/libr2_$1.so pbc_b200/csrc/engine.cu -lcudart & }
build default ""
build fsnolock "-DPBC_FS_LOCKSTEP=0"
wait
ls -la build/libr2_*.so
