#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe_ps.log 2>&1; echo "probe rc=$?"
PBC_B200_LIB=$PWD/gpurun_out/libpbc_b200_os.so timeout 600 python tools/gpu_probe.py > gpurun_out/probe_os.log 2>&1; echo "probe_os rc=$?"
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/probe_ps.log; echo ---- OS; grep pairing_a gpurun_out/probe_os.log
