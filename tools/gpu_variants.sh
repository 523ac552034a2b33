#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for v in $VARIANTS; do
  PBC_B200_LIB=$PWD/build/lib$v.so timeout 600 python tools/gpu_variant_probe.py $WL >> gpurun_out/variants.jsonl 2>gpurun_out/variants_$v.err || echo "variant $v failed: $(tail -2 gpurun_out/variants_$v.err)"
done
cat gpurun_out/variants.jsonl
