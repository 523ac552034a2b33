#!/bin/bash
# parity pass on the GPU box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
