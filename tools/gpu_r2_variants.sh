#!/bin/bash
mkdir -p gpurun_out
VARIANTS="${VARIANTS:-r2_default r2_os}" WL="${WL:-a f 524288}" bash tools/gpu_variants.sh > /dev/null; cp gpurun_out/variants.jsonl gpurun_out/r2_variants_${TAG:-os}.jsonl
cut -c1-330 gpurun_out/r2_variants_${TAG:-os}.jsonl
