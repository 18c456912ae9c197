#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m "gpu and not multigpu" -x -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c13_bench.json')); print(d['value'], 'e2e', d['e2e']['value'], 'agg_us', d['agg_bcast_us_per_round'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c13_trace.txt > gpurun_out/r2c13_trace.log 2>&1; echo "rc=$?"; head -8 gpurun_out/r2c13_trace.txt
