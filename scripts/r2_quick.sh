#!/usr/bin/env bash
# quick A/B session: targeted tests, bench with / without a switch, in-graph timeline with intra-kernel points
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('kernels_per_local_step'), d.get('final_loss'))"; }
timeout 300 python -m pytest tests -m gpu -q --tb=short -k "stem or head or explicit or resnet or maxpool" 2>&1 | grep -v "Warning\|warn" | tail -40
echo "== default"; b
echo "== head wgrad in-kernel"; BATON_HEAD_WGRAD_BRANCH=0 b
echo "== last wgrad cap 64"; BATON_WGRAD_LAST_MAX_CTAS=64 b
echo "== stem fused off"; BATON_STEM_FUSED=0 b
echo "== sgd overlap layer1 split, 148 ctas"; BATON_SGD_OVERLAP=1 b
echo "== sgd overlap layer1 split, 296 ctas"; BATON_SGD_OVERLAP=1 BATON_SGD_TAIL_CTAS=296 b
echo "== sgd overlap layer1 split, 592 ctas"; BATON_SGD_OVERLAP=1 BATON_SGD_TAIL_CTAS=592 b
BATON_TRACE=1 timeout 200 python scripts/trace_step.py --points linear_xent_head_kernel > gpurun_out/r2_trace_quick.txt 2>&1
grep -A12 "intra-kernel points" gpurun_out/r2_trace_quick.txt | head -14
