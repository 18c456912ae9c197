#!/usr/bin/env bash
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
echo "== mp_api_check"; run 240 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT|diag" gpurun_out/r2_api_check_${N}gpu.txt | tail -20
echo "== mp_fedavg_check"; run 200 tests/mp_fedavg_check.py > gpurun_out/r2_fedavg_check_${N}gpu.txt 2>&1; grep -E "FAIL|RESULT|Error" gpurun_out/r2_fedavg_check_${N}gpu.txt | tail -8
echo "== ours engine (prepack + side-stream collective)"; run 200 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours_v3.json 2> gpurun_out/r2_bench_${N}gpu_ours_v3.err; cut -c1-330 gpurun_out/r2_bench_${N}gpu_ours_v3.json; echo; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu_ours_v3.json')); print('e2e', d['e2e']['value'], 'agg_us', d['agg_bcast_us_per_round'], 'loss', d['final_loss'])"
BATON_PREPACK=0 BATON_COLLECTIVE_OVERLAP=0 run 200 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours_v3_noprepack.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu_ours_v3_noprepack.json')); print('no prepack/overlap: value', d['value'], 'e2e', d['e2e']['value'], 'agg_us', d['agg_bcast_us_per_round'], 'loss', d['final_loss'])"
