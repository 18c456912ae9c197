"""Run-to-run spread of the MXFP8 ResNet-18 training smoke test (tests/test_gpu_fp8.py): prints loss histories."""
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
for cfg in (dict(n_epoch=6, lr=0.05, momentum=0.9), dict(n_epoch=8, lr=0.02, momentum=0.9), dict(n_epoch=8, lr=0.05, momentum=0.0)):
    for rep in range(6):
        torch.manual_seed(0)
        X, y = image_shard(ShardSpec(0, torch.full((10,), 0.1), 512), noise=0.3)
        X, y = X.to(dev).to(BF16), y.to(dev)
        m = resnet18(10).set_precision("fp8")
        arena = ParamArena(m, dev, momentum=True)
        m.build_workspace(dev)
        m._graphed_trainer = GraphedLocalSGD(m, arena, loss="ce")
        hist = m.train(X, y, batch_size=128, **cfg)
        print(cfg, ["{:.3f}".format(h) for h in hist], "ratio {:.3f}".format(hist[-1] / hist[0]), flush=True)
