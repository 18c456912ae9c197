#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not multigpu" -x -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -6
BATON_TRACE=1 timeout 200 python scripts/trace_gemm_anatomy.py 2>&1 | grep -v -i warn | tee gpurun_out/r2c16_anatomy.txt | tail -24
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c16_bench.json')); print('r18', d['value'], 'e2e', d['e2e']['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c16_trace.txt > gpurun_out/r2c16_trace.log 2>&1; head -8 gpurun_out/r2c16_trace.txt
timeout 300 python scripts/mb_layers.py > gpurun_out/r2c16_mb_layers.txt 2>&1; cut -c1-220 gpurun_out/r2c16_mb_layers.txt | tail -8
timeout 300 python scripts/microbench.py > gpurun_out/r2c16_mb.log 2>&1; tail -32 gpurun_out/r2c16_mb.log
