#!/usr/bin/env bash
# round-2 first GPU call: state of HEAD + experimental kernels, each under its own timeout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2c1_smi.txt
bash scripts/validate_experimental.sh 2>&1 | tee gpurun_out/r2c1_validate.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"; cat gpurun_out/r2c1_bench.json
BATON_BN_BWD_FUSED=1 BATON_CONV_IGEMM=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench_exp.json 2> gpurun_out/r2c1_bench_exp.err; echo "bench exp rc=$?"; cat gpurun_out/r2c1_bench_exp.json; tail -5 gpurun_out/r2c1_bench_exp.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/r2c1_launches.csv python scripts/profile_step.py --steps 6 > gpurun_out/r2c1_prof.log 2>&1; echo "ncu rc=$?"
