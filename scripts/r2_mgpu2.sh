#!/usr/bin/env bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
echo "== mp_fedavg_check ($N ranks, cooperative launch, no closing barrier)"; run 200 tests/mp_fedavg_check.py > gpurun_out/r2_fedavg_check_${N}gpu.txt 2>&1; grep -E "FAIL|RESULT|Error" gpurun_out/r2_fedavg_check_${N}gpu.txt | tail -8
echo "== mp_api_check"; run 240 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT|Error|error" gpurun_out/r2_api_check_${N}gpu.txt | tail -20
echo "== ours api http"; run 240 bench.py --api http --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_api_http.json 2> gpurun_out/r2_bench_${N}gpu_api_http.err; cat gpurun_out/r2_bench_${N}gpu_api_http.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_api_http.err | tail -5
echo "== ours engine"; run 200 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours_v2.json 2> gpurun_out/r2_bench_${N}gpu_ours_v2.err; cut -c1-330 gpurun_out/r2_bench_${N}gpu_ours_v2.json; echo
echo "== agg bench (trace build: phase stamps)"; BATON_TRACE=1 AGG_PHASES=1 AGG_CTAS=148 AGG_MODELS=resnet18 AGG_WIRES=bf16 run 200 scripts/agg_bench.py 2>&1 | grep -v -i warn | tail -12
