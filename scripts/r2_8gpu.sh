#!/usr/bin/env bash
# The round-2 8-GPU session: correctness logs, flagship + API + baselines, collective phases, BASELINE configs 3/4/5.
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
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
echo "== mp_fedavg_check ($N ranks)"; run 240 tests/mp_fedavg_check.py > gpurun_out/r2_fedavg_check_${N}gpu.txt 2>&1; grep -E "FAIL|RESULT|Error" gpurun_out/r2_fedavg_check_${N}gpu.txt | tail -6
echo "== mp_api_check ($N ranks)"; run 300 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT" gpurun_out/r2_api_check_${N}gpu.txt | tail -12
echo "== ours engine"; run 240 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours.json 2> gpurun_out/r2_bench_${N}gpu_ours.err; show gpurun_out/r2_bench_${N}gpu_ours.json
echo "== ours engine, bcast_gemm"; run 240 bench.py --gpus $N --steps 10 --warmup 3 --bcast-gemm 1 > gpurun_out/r2_bench_${N}gpu_ours_k3.json 2> gpurun_out/r2_bench_${N}gpu_ours_k3.err; show gpurun_out/r2_bench_${N}gpu_ours_k3.json
echo "== ours api http"; run 300 bench.py --api http --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_api_http.json 2> gpurun_out/r2_bench_${N}gpu_api_http.err; show gpurun_out/r2_bench_${N}gpu_api_http.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_api_http.err | grep -i -E "error|Traceback" | head -3
echo "== baseline graph"; run 300 bench.py --impl baseline --graph --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_baseline_graph.json 2> gpurun_out/r2_bench_${N}gpu_baseline_graph.err; show gpurun_out/r2_bench_${N}gpu_baseline_graph.json
echo "== baseline eager"; run 300 bench.py --impl baseline --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_baseline_eager.json 2> gpurun_out/r2_bench_${N}gpu_baseline_eager.err; show gpurun_out/r2_bench_${N}gpu_baseline_eager.json
echo "== agg bench with phase stamps"; BATON_TRACE=1 AGG_PHASES=1 AGG_CTAS=148 AGG_MODELS=resnet18,bert_base AGG_WIRES=bf16,fp8 run 300 scripts/agg_bench.py 2>&1 | grep -v -i "warn\|OMP_NUM\|\*\*\*" | tee gpurun_out/r2_agg_bench_${N}gpu.txt | tail -24
echo "== bert_base local_epochs=5 batch 32"; run 400 bench.py --gpus $N --model bert_base --local-epochs 5 --batch-size 32 --samples 1024 --lr 0.01 --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu_bert.json 2> gpurun_out/r2_bench_${N}gpu_bert.err; show gpurun_out/r2_bench_${N}gpu_bert.json
echo "== resnet50 fp8 alpha 0.1"; run 400 bench.py --gpus $N --model resnet50 --dtype fp8 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50fp8.json 2> gpurun_out/r2_bench_${N}gpu_r50fp8.err; show gpurun_out/r2_bench_${N}gpu_r50fp8.json
echo "== resnet50 bf16 alpha 0.1"; run 400 bench.py --gpus $N --model resnet50 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50bf16.json 2> gpurun_out/r2_bench_${N}gpu_r50bf16.err; show gpurun_out/r2_bench_${N}gpu_r50bf16.json
echo "== sampling 16 logical / 4 sampled"; run 300 bench.py --gpus $N --steps 10 --warmup 3 --logical-clients 16 --sample-k 4 > gpurun_out/r2_bench_${N}gpu_sampling.json 2> gpurun_out/r2_bench_${N}gpu_sampling.err; show gpurun_out/r2_bench_${N}gpu_sampling.json
echo "== sampling 16 logical / 16 (all)"; run 300 bench.py --gpus $N --steps 5 --warmup 3 --logical-clients 16 > gpurun_out/r2_bench_${N}gpu_logical16.json 2> gpurun_out/r2_bench_${N}gpu_logical16.err; show gpurun_out/r2_bench_${N}gpu_logical16.json
