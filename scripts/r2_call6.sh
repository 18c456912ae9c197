#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "batchnorm or fused_column or conv_bn_fused or cluster" -p no:cacheprovider 2>&1 | tail -5
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c6_trace.txt > /dev/null 2>&1; echo "rc=$?"; head -12 gpurun_out/r2c6_trace.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2c6_bench.json
