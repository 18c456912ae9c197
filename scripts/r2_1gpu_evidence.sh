#!/usr/bin/env bash
# The round-2 single-GPU evidence session at HEAD: GPU tests, flagship + strong baseline, fused-attention A/B on BERT,
# ncu --set full of one whole local step + the collective, launch list, compute-sanitizer, in-graph timeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print('  ', sys.argv[1], round(d['value']), d['unit'], 'ms/round', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'agg_us', d.get('agg_bcast_us_per_round'), 'k/step', d.get('kernels_per_local_step'), 'loss', d.get('final_loss'))" $1 2>/dev/null || echo "   no result in $1"; }
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_ours.json 2> gpurun_out/r2_bench_1gpu_ours.err; show gpurun_out/r2_bench_1gpu_ours.json
timeout 300 python bench.py --impl baseline --graph --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu_baseline_graph.json 2> gpurun_out/r2_bench_1gpu_baseline_graph.err; show gpurun_out/r2_bench_1gpu_baseline_graph.json
echo "== bert_base batch 128 x seq 128, fused attention off / on"
BATON_FUSED_ATTN=0 timeout 300 python bench.py --model bert_base --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_bert_attn0.json 2> gpurun_out/r2_bench_1gpu_bert_attn0.err; show gpurun_out/r2_bench_1gpu_bert_attn0.json
timeout 300 python bench.py --model bert_base --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_bert_attn1.json 2> gpurun_out/r2_bench_1gpu_bert_attn1.err; show gpurun_out/r2_bench_1gpu_bert_attn1.json
echo "== ncu --set full: one whole eager local step (warm caches), summarised on the box; then the collective + optimizer"
# the full-step report is ~150 MB: it stays on the box (gpurun_out/ is capped at 64 MiB), only its text summary travels
timeout 500 ncu --set full --clock-control none --cache-control none -s 300 -c 125 -o /tmp/r2_ncu_step python scripts/profile_step.py --steps 4 --agg 0 > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu step rc=$?"
python scripts/ncu_summary.py /tmp/r2_ncu_step.ncu-rep > gpurun_out/r2_ncu_step_summary.txt 2>&1
timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"gemm_bf16_fixed_kernel|bn_bwd_cluster_kernel|bn_apply_kernel|gemm_bf16_tcgen05_kernel|linear_xent_head" -s 150 -c 10 -o gpurun_out/r2_ncu_top python scripts/profile_step.py --steps 4 --agg 0 > gpurun_out/r2_ncu_top.log 2>&1; echo "ncu top rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"fedavg|fused_sgd" -c 4 -o gpurun_out/r2_ncu_fedavg python scripts/profile_step.py --steps 1 --agg 2 > gpurun_out/r2_ncu_fedavg.log 2>&1; echo "ncu fedavg rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 360 --csv --log-file gpurun_out/r2_launches.csv python scripts/profile_step.py --steps 6 --agg 1 > gpurun_out/r2_prof.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3; du -sh gpurun_out
echo "== sanitizers"; bash scripts/r2_sanitize.sh
echo "== in-graph timeline (trace build)"
BATON_TRACE=1 timeout 200 python scripts/trace_step.py > gpurun_out/r2_trace_resnet18_head.txt 2>&1; head -22 gpurun_out/r2_trace_resnet18_head.txt | grep -v -i warn
timeout 200 python bench.py --impl baseline --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_baseline_eager.json 2> gpurun_out/r2_bench_1gpu_baseline_eager.err; show gpurun_out/r2_bench_1gpu_baseline_eager.json
