#!/usr/bin/env bash
# compute-sanitizer over every kernel family (single GPU) and memcheck over the 2-rank collective check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool python scripts/sanitize_smoke.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "[$tool] rc=$? $(grep -E 'sanitize smoke done' gpurun_out/r2_sanitizer_$tool.log | tail -1) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_$tool.log | tail -1)"
done
