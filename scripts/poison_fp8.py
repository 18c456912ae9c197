"""Uninitialised-read hunt: fill the caching allocator's free blocks with NaN bit patterns, then run the MXFP8 (and the
bf16) ResNet-18 training smoke.  A kernel that consumes a torch.empty() buffer it did not fully write turns the loss
into NaN here, while it passes on a fresh process (fresh device memory reads as zeros)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.data import ShardSpec, image_shard  # noqa: E402
from baton_b200.models import resnet18  # noqa: E402
from baton_b200.parallel.arena import ParamArena  # noqa: E402
from baton_b200.train import GraphedLocalSGD  # noqa: E402

dev = torch.device("cuda:0")
BF16 = torch.bfloat16


def poison():
    blocks = [torch.full((n,), -1, dtype=torch.int32, device=dev) for n in (1 << 28, 1 << 26, 1 << 24, 1 << 22) for _ in range(3)]
    blocks += [torch.full((1 << k,), -1, dtype=torch.int32, device=dev) for k in range(8, 22) for _ in range(16)]
    torch.cuda.synchronize()
    del blocks


for prec in (sys.argv[1:] or ["fp8", "bf16"]):
    for use_graph in (False, True):
        poison()
        torch.manual_seed(0)
        X, y = image_shard(ShardSpec(0, torch.full((10,), 0.1), 512), noise=0.3)
        X, y = X.to(dev).to(BF16), y.to(dev)
        m = resnet18(10)
        if prec == "fp8":
            m.set_precision("fp8")
        arena = ParamArena(m, dev, momentum=True)
        m.build_workspace(dev)
        m._graphed_trainer = GraphedLocalSGD(m, arena, loss="ce", use_graph=use_graph)
        hist = m.train(X, y, n_epoch=3, lr=0.05, batch_size=128, momentum=0.9)
        print(prec, "graph" if use_graph else "eager", ["{:.3f}".format(h) for h in hist], flush=True)
