#!/bin/bash
# round 2, first GPU call: the A/B builds DESIGN.md section 8 queued (A1 occupancy + NAF, NAF scan of r for F/D/G)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_first_smi.csv
VARIANTS="a1_default a1_s13 a1_naf a1_s13naf" WL="a1 28416" bash tools/gpu_variants.sh > /dev/null
cp gpurun_out/variants.jsonl gpurun_out/r2_variants_a1.jsonl
VARIANTS="a1_default cc_naf" WL="f d g 131072" bash tools/gpu_variants.sh > /dev/null
cp gpurun_out/variants.jsonl gpurun_out/r2_variants_cc_naf.jsonl
cut -c1-400 gpurun_out/r2_variants_a1.jsonl gpurun_out/r2_variants_cc_naf.jsonl
for v in a1_s13naf; do PBC_B200_LIB=$PWD/build/lib$v.so timeout 900 python -m pytest tests/test_gpu_type_a1.py -m gpu -q 2>&1 | tail -2; done
PBC_B200_LIB=$PWD/build/libcc_naf.so timeout 900 python -m pytest tests/test_gpu_type_fd.py -m gpu -q 2>&1 | tail -2
