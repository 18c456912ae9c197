#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "explicit or maxpool or graphed or fused_sgd" 2>&1 | grep -v Warning | tail -12
for v in "BATON_SGD_OVERLAP=0" "BATON_SGD_TAIL_CTAS=148" "BATON_SGD_TAIL_CTAS=64" "BATON_SGD_TAIL_CTAS=32"; do
  env $v BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c10_trace_$v.txt > gpurun_out/r2c10_trace.log 2>&1; echo "$v rc=$?"; head -1 gpurun_out/r2c10_trace_$v.txt; tail -2 gpurun_out/r2c10_trace.log | cut -c1-200
done
