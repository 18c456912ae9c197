#!/usr/bin/env bash
# 2-GPU evidence at HEAD: flagship with the in-repo NVLink measurement, client sampling end to end, compute-sanitizer
# memcheck over the 2-rank collective check.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { port=$((29600 + RANDOM % 300)); timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "${@:2}"; }
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("   value {:.0f} ms/round {:.2f} e2e {:.0f} agg_us {} roofline {} link {} loss {}".format(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("agg_bcast_us_per_round"), d.get("agg_bcast_roofline"), d.get("nvlink_GBps_per_dir_measured_here"), d.get("final_loss")))
except Exception as e:
    print("   no result:", e)
PY
}
echo "== ours"; run 200 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_ours.json 2> gpurun_out/r2_bench_${N}gpu_ours.err; show gpurun_out/r2_bench_${N}gpu_ours.json
echo "== sampling: $((2*N)) logical / $N sampled"; run 200 bench.py --gpus $N --steps 10 --warmup 3 --logical-clients $((2*N)) --sample-k $N > gpurun_out/r2_bench_${N}gpu_sampling.json 2> gpurun_out/r2_bench_${N}gpu_sampling.err; show gpurun_out/r2_bench_${N}gpu_sampling.json
echo "== mp_fedavg_check"; run 200 tests/mp_fedavg_check.py > gpurun_out/r2_fedavg_check_${N}gpu.txt 2>&1; grep -E "FAIL|RESULT" gpurun_out/r2_fedavg_check_${N}gpu.txt | tail -4
echo "== compute-sanitizer memcheck over the $N-rank collective check"
port=$((29600 + RANDOM % 300))
BATON_CHECK_SKIP_DEAD_PEER=1 timeout 280 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r2_sanitizer_mp_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/mp_fedavg_check.py > gpurun_out/r2_sanitizer_mp_stdout.txt 2>&1
echo "rc=$?"; grep -E "FAIL|RESULT" gpurun_out/r2_sanitizer_mp_stdout.txt | tail -3; grep -h "ERROR SUMMARY" gpurun_out/r2_sanitizer_mp_*.log | sort | uniq -c
