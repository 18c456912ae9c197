#!/usr/bin/env bash
# The round-2 multi-GPU session (N = 8 by default): correctness logs, flagship + API + baselines, collective phases,
# BASELINE configs 3/4/5.  Steps run in order of importance; a global time budget skips what does not fit.
N=${1:-8}
LIMIT=${2:-540}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
left() { echo $(( LIMIT - ( $(date +%s) - T0 ) )); }
run() { # timeout, args...
  local t=$1; shift
  local l=$(left)
  if [ "$l" -lt 40 ]; then echo "   SKIPPED (time budget)"; return 124; fi
  if [ "$t" -gt "$l" ]; then t=$l; fi
  port=$((29600 + RANDOM % 300)); timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   value {:.0f} {} ms/round {:.2f} e2e {:.0f} agg_us {} roofline {} loss {} cfg {}".format(d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("agg_bcast_us_per_round"), (d.get("agg_bcast_roofline") or {}).get("fraction_of_measured"), d.get("final_loss"), {k: d["config"].get(k) for k in ("model", "local_epochs", "batch_size", "samples_per_client", "logical_clients", "sampled_per_round", "wire_dtype", "nvls_choice", "cuda_graph")}))
    for k in ("control_plane_ms_per_round", "local_train_ms_per_round", "replicas_identical"):
        if k in d: print("   ", k, d[k])
except Exception as e:
    print("   no result:", e)
PY
}
echo "== mp_fedavg_check ($N ranks)"; run 150 tests/mp_fedavg_check.py > gpurun_out/r2_fedavg_check_${N}gpu.txt 2>&1; grep -E "FAIL|RESULT|Error" gpurun_out/r2_fedavg_check_${N}gpu.txt | tail -6
echo "== ours engine [$(left)s left]"; run 150 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours.json 2> gpurun_out/r2_bench_${N}gpu_ours.err; show gpurun_out/r2_bench_${N}gpu_ours.json
echo "== ours api http [$(left)s left]"; run 200 bench.py --api http --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_api_http.json 2> gpurun_out/r2_bench_${N}gpu_api_http.err; show gpurun_out/r2_bench_${N}gpu_api_http.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_api_http.err | grep -i -E "error|Traceback" | head -3
echo "== agg bench with phase stamps [$(left)s left]"; BATON_TRACE=1 AGG_PHASES=1 AGG_CTAS=148 AGG_MODELS=resnet18,bert_base AGG_WIRES=bf16,fp8 run 200 scripts/agg_bench.py 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*" | tee gpurun_out/r2_agg_bench_${N}gpu.txt | tail -24
echo "== bert_base local_epochs=5 batch 32 [$(left)s left]"; run 250 bench.py --gpus $N --model bert_base --local-epochs 5 --batch-size 32 --samples 1024 --lr 0.01 --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu_bert.json 2> gpurun_out/r2_bench_${N}gpu_bert.err; show gpurun_out/r2_bench_${N}gpu_bert.json
echo "== sampling 16 logical / 4 sampled [$(left)s left]"; run 150 bench.py --gpus $N --steps 10 --warmup 3 --logical-clients 16 --sample-k 4 > gpurun_out/r2_bench_${N}gpu_sampling.json 2> gpurun_out/r2_bench_${N}gpu_sampling.err; show gpurun_out/r2_bench_${N}gpu_sampling.json
echo "== resnet50 fp8 alpha 0.1 [$(left)s left]"; run 200 bench.py --gpus $N --model resnet50 --dtype fp8 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50fp8.json 2> gpurun_out/r2_bench_${N}gpu_r50fp8.err; show gpurun_out/r2_bench_${N}gpu_r50fp8.json
echo "== baseline graph [$(left)s left]"; run 200 bench.py --impl baseline --graph --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_baseline_graph.json 2> gpurun_out/r2_bench_${N}gpu_baseline_graph.err; show gpurun_out/r2_bench_${N}gpu_baseline_graph.json
echo "== mp_api_check ($N ranks) [$(left)s left]"; run 200 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT" gpurun_out/r2_api_check_${N}gpu.txt | tail -12
echo "== ours engine, bcast_gemm [$(left)s left]"; run 150 bench.py --gpus $N --steps 10 --warmup 3 --bcast-gemm 1 > gpurun_out/r2_bench_${N}gpu_ours_k3.json 2> gpurun_out/r2_bench_${N}gpu_ours_k3.err; show gpurun_out/r2_bench_${N}gpu_ours_k3.json
echo "== resnet50 bf16 alpha 0.1 [$(left)s left]"; run 200 bench.py --gpus $N --model resnet50 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50bf16.json 2> gpurun_out/r2_bench_${N}gpu_r50bf16.err; show gpurun_out/r2_bench_${N}gpu_r50bf16.json
echo "== baseline eager [$(left)s left]"; run 200 bench.py --impl baseline --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_baseline_eager.json 2> gpurun_out/r2_bench_${N}gpu_baseline_eager.err; show gpurun_out/r2_bench_${N}gpu_baseline_eager.json
echo "== 16 logical clients, all sampled [$(left)s left]"; run 150 bench.py --gpus $N --steps 5 --warmup 3 --logical-clients 16 > gpurun_out/r2_bench_${N}gpu_logical16.json 2> gpurun_out/r2_bench_${N}gpu_logical16.err; show gpurun_out/r2_bench_${N}gpu_logical16.json
echo "== done in $(( $(date +%s) - T0 )) s"
