#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# cluster split-K kernel in situ (eager steps, warm caches): why 16 us in-graph vs 7 us back-to-back?
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 40 -c 4 \
  -o gpurun_out/r2c3_cluster python scripts/profile_step.py --steps 4 --agg 0 > gpurun_out/r2c3_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python scripts/microbench.py > gpurun_out/r2c3_mb.log 2>&1; echo "mb rc=$?"; tail -30 gpurun_out/r2c3_mb.log
