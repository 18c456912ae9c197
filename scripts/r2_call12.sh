#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not multigpu" -x -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c12_bench.json')); print(d['value'], 'e2e', d['e2e']['value'], 'agg_us', d['agg_bcast_us_per_round'], 'k/step', d['kernels_per_local_step'], d['config'].get('bcast_gemm'), d['config'].get('upload_copy_emitted_by_sgd'), 'loss', d['final_loss'])"; tail -3 gpurun_out/r2c12_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --bcast-gemm 1 > gpurun_out/r2c12_bench_k3.json 2> gpurun_out/r2c12_bench_k3.err; echo "bench k3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c12_bench_k3.json')); print(d['value'], 'e2e', d['e2e']['value'], 'agg_us', d['agg_bcast_us_per_round'], 'k/step', d['kernels_per_local_step'], d['config'].get('bcast_gemm'), d['config'].get('upload_copy_emitted_by_sgd'), 'loss', d['final_loss'])"; tail -3 gpurun_out/r2c12_bench_k3.err
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c12_trace.txt > gpurun_out/r2c12_trace.log 2>&1; echo "rc=$?"; head -12 gpurun_out/r2c12_trace.txt
for v in "BATON_BN_BWD_MAX_CLUSTER=8" "BATON_BN_BWD_MAX_CLUSTER=4" "BATON_WGRAD_OVERLAP=0"; do
  env $v BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out "gpurun_out/r2c12_trace_$v.txt" > gpurun_out/r2c12_trace.log 2>&1; echo "$v rc=$?"; head -1 "gpurun_out/r2c12_trace_$v.txt"
done
