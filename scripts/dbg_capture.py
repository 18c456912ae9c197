"""Regression probe: a model used eagerly on the default stream must still be CUDA-graph capturable."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.models import bert_tiny  # noqa: E402
from baton_b200.ops import nn as bnn  # noqa: E402
from baton_b200.parallel.arena import ParamArena  # noqa: E402
from baton_b200.train import GraphedLocalSGD  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = bert_tiny(3)
arena = ParamArena(m, dev)
ids = torch.randint(0, 1024, (8, 64), device=dev)
y = torch.randint(0, 3, (8,), device=dev)
loss, _ = bnn.cross_entropy(m(ids), y)
loss.backward()
arena.grad.zero_()
X = torch.randint(0, 1024, (256, 64), device=dev)
yy = (X[:, :8].sum(1) % 3)
tr = GraphedLocalSGD(m, arena, loss="ce")
m._graphed_trainer = tr
print("OK", [round(h, 3) for h in m.train(X, yy, n_epoch=3, lr=0.05, batch_size=32)])
