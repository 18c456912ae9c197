#!/usr/bin/env bash
# One multi-GPU box session (N = $1): collective correctness log, ours / strong baseline / API-over-HTTP benches.
N=${1:-2}
TAG=${2:-r2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
echo "== mp_fedavg_check ($N ranks)"; run 200 tests/mp_fedavg_check.py > gpurun_out/${TAG}_fedavg_check_${N}gpu.txt 2>&1; grep -E "PASS|FAIL|RESULT|Error" gpurun_out/${TAG}_fedavg_check_${N}gpu.txt | tail -25
echo "== ours engine"; run 200 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_ours.json 2> gpurun_out/${TAG}_bench_${N}gpu_ours.err; cut -c1-330 gpurun_out/${TAG}_bench_${N}gpu_ours.json; echo
echo "== ours api http"; run 300 bench.py --api http --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_api_http.json 2> gpurun_out/${TAG}_bench_${N}gpu_api_http.err; cat gpurun_out/${TAG}_bench_${N}gpu_api_http.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu_api_http.err
echo "== baseline graph"; run 300 bench.py --impl baseline --graph --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_baseline_graph.json 2> gpurun_out/${TAG}_bench_${N}gpu_baseline_graph.err; cut -c1-330 gpurun_out/${TAG}_bench_${N}gpu_baseline_graph.json; tail -3 gpurun_out/${TAG}_bench_${N}gpu_baseline_graph.err
echo "== baseline eager"; run 300 bench.py --impl baseline --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_baseline_eager.json 2> gpurun_out/${TAG}_bench_${N}gpu_baseline_eager.err; cut -c1-330 gpurun_out/${TAG}_bench_${N}gpu_baseline_eager.json
echo "== multigpu pytest"; timeout 400 python -m pytest tests -m "gpu and multigpu" -x -q -p no:cacheprovider 2>&1 | grep -v Warning | tail -5
