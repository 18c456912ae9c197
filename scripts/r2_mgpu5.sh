#!/usr/bin/env bash
# validation of the remaining multi-GPU commands at N GPUs before the 8-GPU session
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   value {:.0f} {} ms/round {:.2f} e2e {:.0f} agg_us {} loss {} cfg {}".format(d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("agg_bcast_us_per_round"), d.get("final_loss"), {k: d["config"].get(k) for k in ("model", "local_epochs", "batch_size", "samples_per_client", "logical_clients", "sampled_per_round", "wire_dtype")}))
except Exception as e:
    print("   no result:", e)
PY
}
echo "== mp_api_check"; run 240 tests/mp_api_check.py > gpurun_out/r2_api_check_${N}gpu.txt 2>&1; grep -E "^ok|FAIL|RESULT|diag" gpurun_out/r2_api_check_${N}gpu.txt | tail -12
echo "== bert_base local_epochs=5 batch 32"; run 400 bench.py --gpus $N --model bert_base --local-epochs 5 --batch-size 32 --samples 1024 --lr 0.01 --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu_bert.json 2> gpurun_out/r2_bench_${N}gpu_bert.err; show gpurun_out/r2_bench_${N}gpu_bert.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_bert.err | tail -3
echo "== resnet50 fp8 alpha 0.1"; run 400 bench.py --gpus $N --model resnet50 --dtype fp8 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50fp8.json 2> gpurun_out/r2_bench_${N}gpu_r50fp8.err; show gpurun_out/r2_bench_${N}gpu_r50fp8.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_r50fp8.err | tail -3
echo "== resnet50 bf16 alpha 0.1"; run 400 bench.py --gpus $N --model resnet50 --alpha 0.1 --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu_r50bf16.json 2> gpurun_out/r2_bench_${N}gpu_r50bf16.err; show gpurun_out/r2_bench_${N}gpu_r50bf16.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_r50bf16.err | tail -3
echo "== sampling: $((2*N)) logical / $((N/2 > 0 ? N/2 : 1)) sampled"; run 300 bench.py --gpus $N --steps 10 --warmup 3 --logical-clients $((2*N)) --sample-k $((N/2 > 0 ? N/2 : 1)) > gpurun_out/r2_bench_${N}gpu_sampling.json 2> gpurun_out/r2_bench_${N}gpu_sampling.err; show gpurun_out/r2_bench_${N}gpu_sampling.json; grep -v -i warn gpurun_out/r2_bench_${N}gpu_sampling.err | tail -3
