#!/bin/bash
# compute-sanitizer memcheck + synccheck over one small batch per parameter set; logs -> gpurun_out/
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python tools/gpu_sanitize.py ${SAN_TYPES:-a f d159 g149} > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|sanitize batch" gpurun_out/r2_sanitizer_$tool.log | tail -8
done
