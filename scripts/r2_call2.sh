#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BATON_TRACE=1 timeout 300 python scripts/trace_step.py --model resnet18 --out gpurun_out/r2c2_trace_r18.txt > gpurun_out/r2c2_trace.log 2>&1; echo "trace rc=$?"; tail -5 gpurun_out/r2c2_trace.log
BATON_TRACE=1 BATON_BN_BWD_FUSED=1 BATON_CONV_IGEMM=1 timeout 300 python scripts/trace_step.py --model resnet18 --out gpurun_out/r2c2_trace_r18_exp.txt > gpurun_out/r2c2_trace_exp.log 2>&1; echo "trace exp rc=$?"; tail -3 gpurun_out/r2c2_trace_exp.log
timeout 300 python scripts/microbench.py > gpurun_out/r2c2_mb.log 2>&1; echo "mb rc=$?"
