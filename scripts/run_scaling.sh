#!/usr/bin/env bash
# bench.py at N GPUs for the three arms; results under gpurun_out/scale_<arm>_<N>.json
N=${1:-8}; STEPS=${2:-10}; ARMS=${3:-"ours baseline reference"}
mkdir -p gpurun_out
for arm in $ARMS; do
  port=$((29600 + RANDOM % 300))
  steps=$STEPS; [ "$arm" = reference ] && steps=3
  if [ "$N" = 1 ]; then
    timeout 400 python bench.py --impl $arm --gpus 1 --steps $steps --warmup 3 > gpurun_out/scale_${arm}_${N}.json 2> gpurun_out/scale_${arm}_${N}.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --impl $arm --gpus $N --steps $steps --warmup 3 > gpurun_out/scale_${arm}_${N}.json 2> gpurun_out/scale_${arm}_${N}.err
  fi
  echo "== $arm N=$N rc=$?"; tail -c 1500 gpurun_out/scale_${arm}_${N}.json; echo; tail -n 4 gpurun_out/scale_${arm}_${N}.err
done
