#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c7_trace.txt > gpurun_out/r2c7_trace.log 2>&1; echo "rc=$?"; head -14 gpurun_out/r2c7_trace.txt; tail -3 gpurun_out/r2c7_trace.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c7_bench.json; tail -3 gpurun_out/r2c7_bench.err
