#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print('  ', sys.argv[1], round(d['value']), d['unit'], 'ms/round', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'agg_us', d.get('agg_bcast_us_per_round'), 'k/step', d.get('kernels_per_local_step'), 'loss', d.get('final_loss'))" $1; }
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_ours.json 2> gpurun_out/r2_bench_1gpu_ours.err; show gpurun_out/r2_bench_1gpu_ours.json
timeout 300 python bench.py --impl baseline --graph --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_baseline_graph.json 2> gpurun_out/r2_bench_1gpu_baseline_graph.err; show gpurun_out/r2_bench_1gpu_baseline_graph.json
timeout 300 python bench.py --impl baseline --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu_baseline_eager.json 2> gpurun_out/r2_bench_1gpu_baseline_eager.err; show gpurun_out/r2_bench_1gpu_baseline_eager.json
echo "== bert_base batch 128 x seq 128, fused attention off / on"
BATON_FUSED_ATTN=0 timeout 400 python bench.py --model bert_base --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_bert_attn0.json 2> gpurun_out/r2_bench_1gpu_bert_attn0.err; show gpurun_out/r2_bench_1gpu_bert_attn0.json
BATON_FUSED_ATTN=1 timeout 400 python bench.py --model bert_base --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_bert_attn1.json 2> gpurun_out/r2_bench_1gpu_bert_attn1.err; show gpurun_out/r2_bench_1gpu_bert_attn1.json
echo "== bert_base local_epochs=5 batch 32"
BATON_FUSED_ATTN=1 timeout 400 python bench.py --model bert_base --local-epochs 5 --batch-size 32 --samples 1024 --lr 0.01 --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_bert_le5.json 2> gpurun_out/r2_bench_1gpu_bert_le5.err; show gpurun_out/r2_bench_1gpu_bert_le5.json
echo "== sanitizers"; bash scripts/r2_sanitize.sh
echo "== ncu --set full of the top kernels (eager steps, warm caches)"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"gemm_bf16_fixed_kernel|bn_bwd_cluster_kernel|bn_apply_kernel|gemm_bf16_tcgen05_kernel" -s 120 -c 24 -o gpurun_out/r2_ncu_step python scripts/profile_step.py --steps 3 --agg 0 > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fedavg_allreduce|fused_sgd" -c 4 -o gpurun_out/r2_ncu_fedavg python scripts/profile_step.py --steps 1 --agg 2 > gpurun_out/r2_ncu_fedavg.log 2>&1; echo "ncu fedavg rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 360 --csv --log-file gpurun_out/r2_launches.csv python scripts/profile_step.py --steps 6 --agg 1 > gpurun_out/r2_prof.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
