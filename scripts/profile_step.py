"""Run a few EAGER (un-graphed) local-SGD steps of ResNet-18 so `ncu` can attribute device time to
individual kernels.   ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 6"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.data import ShardSpec, image_shard, token_shard  # noqa: E402
from baton_b200.models import bert_base, resnet18, resnet50  # noqa: E402
from baton_b200.parallel.arena import ParamArena  # noqa: E402
from baton_b200.parallel.fedavg import FedAvgSession  # noqa: E402
from baton_b200.train import GraphedLocalSGD  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--batch-size", type=int, default=128)
ap.add_argument("--model", default="resnet18")
ap.add_argument("--agg", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
is_bert = args.model == "bert_base"
m = bert_base(2) if is_bert else (resnet18(10) if args.model == "resnet18" else resnet50(10))
arena = ParamArena(m, dev)
if not is_bert:
    m.build_workspace(dev)
tr = GraphedLocalSGD(m, arena, loss="ce", use_graph=False)
if is_bert:
    X, y = token_shard(ShardSpec(0, torch.full((2,), 0.5), args.batch_size * args.steps), seq_len=128)
else:
    X, y = image_shard(ShardSpec(0, torch.full((10,), 0.1), args.batch_size * args.steps), dtype=torch.bfloat16)
X, y = X.to(dev), y.to(dev)
tr.run(X, y, n_epoch=1, lr=0.05, batch_size=args.batch_size)
sess = FedAvgSession(arena)
for _ in range(args.agg):
    sess.aggregate(my_n=1.0)
torch.cuda.synchronize()
print("done", tr.last_stats)
