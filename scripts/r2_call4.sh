#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "implicit_gemm or conv2d_forward or tma_im2col" -p no:cacheprovider 2>&1 | tail -5
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c4_trace_default.txt > /dev/null 2>&1; echo "rc=$?"; head -1 gpurun_out/r2c4_trace_default.txt
BATON_GEMM_CLUSTER_MIN_KT=0 BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c4_trace_nocluster.txt > /dev/null 2>&1; echo "rc=$?"; head -1 gpurun_out/r2c4_trace_nocluster.txt
BATON_PDL=0 BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c4_trace_nopdl.txt > /dev/null 2>&1; echo "rc=$?"; head -1 gpurun_out/r2c4_trace_nopdl.txt
BATON_WGRAD_OVERLAP=0 BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c4_trace_nowgradoverlap.txt > /dev/null 2>&1; echo "rc=$?"; head -1 gpurun_out/r2c4_trace_nowgradoverlap.txt
