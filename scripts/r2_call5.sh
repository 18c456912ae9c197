#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/mb_layers.py > gpurun_out/r2c5_mb_layers.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2c5_mb_layers.txt | tail -30
