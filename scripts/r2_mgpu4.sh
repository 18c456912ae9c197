#!/usr/bin/env bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
echo "== mp_api_check"; API_CHECK_VERBOSE=1 run 240 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT|diag|\[rank" gpurun_out/r2_api_check_${N}gpu.txt | tail -60
