#!/usr/bin/env bash
# quick A/B session: targeted tests, bench with / without a switch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('kernels_per_local_step'), d.get('final_loss'))"; }
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -x 2>&1 | grep -v "Warning\|warn" | tail -5
echo "== krot 1 (default)"; b
echo "== krot 0"; BATON_GEMM_KROT=0 b
echo "== krot 1"; b
echo "== krot 0"; BATON_GEMM_KROT=0 b
