#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m "gpu and not multigpu" -x -q -p no:cacheprovider -k "head or explicit or graphed or conv2d or im2col" 2>&1 | grep -v Warning | tail -5
BATON_TRACE=1 timeout 200 python scripts/trace_gemm_anatomy.py 2>&1 | grep -v -i warn | tee gpurun_out/r2c15_anatomy.txt | tail -30
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2c15_bench.json')); print('r18', d['value'], 'e2e', d['e2e']['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --out gpurun_out/r2c15_trace.txt > gpurun_out/r2c15_trace.log 2>&1; head -4 gpurun_out/r2c15_trace.txt
