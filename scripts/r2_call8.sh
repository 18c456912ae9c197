#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "explicit or maxpool or conv2d or gemm" 2>&1 | grep -v Warning | tail -40
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c8_trace.txt > gpurun_out/r2c8_trace.log 2>&1; echo "rc=$?"; head -3 gpurun_out/r2c8_trace.txt
