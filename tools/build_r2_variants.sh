#!/bin/bash
# A/B builds for the round-2 kernels (run HERE; the .so files travel with the snapshot):
#   build/libr2_default.so    the configuration before the change under test
#   build/libr2_acc.so        PBC_FQ_ACC=1: unmerged accumulating products + row-wise reduction (type F slot kernels)
#   build/libr2_accsq.so      ... with the column-wise squarer kept (PBC_FS_SQR_OS=0)
#   build/libr2_acc160.so     ... with 160-thread Miller blocks, registers held to two blocks per SM
#   build/libr2_old160.so     the old products with 160-thread Miller blocks
set -e
cd "$(dirname "$0")/.."
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $FLAGS $2 -Xptxas -v -o build/libr2_$1.so pbc_b200/csrc/engine.cu -lcudart 2> build/ptxas_$1.log & }
build default "-DPBC_FQ_ACC=0"
build acc ""
build accsq "-DPBC_FS_SQR_OS=0"
build acc160 "-DPBC_FS_MILLER_BLOCK=160 -DPBC_FS_MILLER_MAXREG=200"
build old160 "-DPBC_FQ_ACC=0 -DPBC_FS_MILLER_BLOCK=160 -DPBC_FS_MILLER_MAXREG=200"
wait
for v in default acc accsq acc160 old160; do echo $v; grep -A2 "k_f_miller_s\|k_f_finalexp_s" build/ptxas_$v.log | grep "registers\|spill"; done
ls -la build/libr2_*.so
