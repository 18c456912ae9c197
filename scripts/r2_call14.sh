#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m "gpu and not multigpu" -x -q -p no:cacheprovider -k "conv2d or im2col" 2>&1 | grep -v Warning | tail -4
for cfg in "--model resnet50 --alpha 0.1" "--model resnet50 --alpha 0.1 --dtype fp8"; do
  timeout 300 python bench.py $cfg --steps 5 --warmup 3 > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err; echo "bench $cfg rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c14_bench.json')); print('  ', d['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"; tail -2 gpurun_out/r2c14_bench.err | grep -v -i warn
done
BATON_EXPLICIT_STEP=0 timeout 300 python bench.py --model resnet50 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err; echo "bench r50 autograd path rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c14_bench.json')); print('  ', d['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
BATON_WGRAD_MAX_CTAS=148 timeout 300 python bench.py --model resnet50 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2c14_bench.json 2> gpurun_out/r2c14_bench.err; echo "bench r50 wgrad cap 148 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c14_bench.json')); print('  ', d['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --model resnet50 --steps 4 --out gpurun_out/r2c14_trace_r50.txt > gpurun_out/r2c14_trace.log 2>&1; echo "rc=$?"; head -16 gpurun_out/r2c14_trace_r50.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c14_bench18.json 2> gpurun_out/r2c14_bench18.err; python -c "
import json; d=json.load(open('gpurun_out/r2c14_bench18.json')); print('r18', d['value'], 'ms/round', d['ms_per_step'], 'k/step', d['kernels_per_local_step'], 'loss', d['final_loss'])"
