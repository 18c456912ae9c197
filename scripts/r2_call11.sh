#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "BATON_WGRAD_MAX_CTAS=148" "BATON_WGRAD_MAX_CTAS=96" "BATON_WGRAD_MAX_CTAS=64" "BATON_WGRAD_MAX_CTAS=32" "BATON_BRANCH_OVERLAP=0"; do
  env $v BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c11_trace_$v.txt > gpurun_out/r2c11_trace.log 2>&1; echo "$v rc=$?"; head -1 gpurun_out/r2c11_trace_$v.txt
done
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2c11_bench.json
timeout 300 python bench.py --impl baseline --steps 10 --warmup 3 > gpurun_out/r2c11_base_eager.json 2> gpurun_out/r2c11_base_eager.err; echo "base eager rc=$?"; cut -c1-330 gpurun_out/r2c11_base_eager.json
timeout 300 python bench.py --impl baseline --graph --steps 10 --warmup 3 > gpurun_out/r2c11_base_graph.json 2> gpurun_out/r2c11_base_graph.err; echo "base graph rc=$?"; cut -c1-330 gpurun_out/r2c11_base_graph.json; tail -5 gpurun_out/r2c11_base_graph.err
