#!/usr/bin/env bash
# One 8-GPU box session: correctness of the fused collective at 8 ranks, ours + reference bench, agg/bcast
# microbenchmark, and the client-sampling config.  Every step is bounded by its own timeout.
N=${1:-8}
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
echo "== mp_fedavg_check"; run 150 tests/mp_fedavg_check.py 2>&1 | grep -E "FAIL|RESULT|Error" | head -10
echo "== ours"; run 150 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_ours_$N.json 2> gpurun_out/scale_ours_$N.err; tail -c 900 gpurun_out/scale_ours_$N.json; echo
echo "== reference"; run 240 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_reference_$N.json 2> gpurun_out/scale_reference_$N.err; head -c 400 gpurun_out/scale_reference_$N.json; echo
echo "== agg"; AGG_CTAS=296 AGG_MODELS=resnet18,bert_base AGG_WIRES=bf16,fp8 run 150 scripts/agg_bench.py 2>&1 | grep -E "fused|nccl" | head -14
echo "== sampling 16 logical / 4 sampled"; run 150 bench.py --gpus $N --steps 10 --warmup 3 --logical-clients 16 --sample-k 4 > gpurun_out/scale_sampling_$N.json 2> gpurun_out/scale_sampling_$N.err; head -c 500 gpurun_out/scale_sampling_$N.json; echo
