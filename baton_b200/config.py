"""Typed configuration for federated jobs.

The reference has no config system: three positional argv (demo.py:64-66) and
constructor defaults -- ``client_ttl=300`` (manager.py:22), ``heartbeat_time=60``,
``port=8080`` (worker.py:13-14), ``n_epoch=32`` (manager.py:55), ``lr=0.001,
batch_size=32`` (demo.py:29).  Those defaults are preserved here and extended
with the knobs the BASELINE.json configs need.
"""
from __future__ import annotations

import argparse
import json
from dataclasses import asdict, dataclass, fields
from typing import Optional


@dataclass
class FederationConfig:
    model: str = "lineartest"          # lineartest | mlp2 | resnet18 | resnet50 | bert_base
    dtype: str = "fp32"                # fp32 | bf16 | fp8 (block-scaled mxfp8 GEMMs)
    backend: str = "http"              # http | fused | nccl
    clients: int = 2                   # physical clients (GPUs)
    logical_clients: int = 0           # >clients => time-sliced logical clients
    sample_k: Optional[int] = None     # participants per round (None = all)
    local_epochs: int = 32             # manager.py:55
    lr: float = 0.001                  # demo.py:29
    batch_size: int = 32               # demo.py:29
    momentum: float = 0.0
    weight_decay: float = 0.0
    partition: str = "iid"             # iid | label_skew | dirichlet
    alpha: float = 0.1                 # Dirichlet concentration
    samples_per_client: int = 4096
    num_classes: int = 10
    client_ttl: float = 300.0          # manager.py:22
    heartbeat_time: float = 60.0       # worker.py:14
    round_timeout: Optional[float] = None
    wire_dtype: str = "bf16"           # precision of the upload over NVLink: fp32 | bf16
    checkpoint_dir: Optional[str] = None
    seed: int = 0

    def to_json(self) -> str:
        return json.dumps(asdict(self), sort_keys=True)

    @classmethod
    def from_json(cls, text: str) -> "FederationConfig":
        known = {f.name for f in fields(cls)}
        data = json.loads(text)
        unknown = set(data) - known
        if unknown:
            raise ValueError("unknown config keys: {}".format(sorted(unknown)))
        return cls(**data)

    @classmethod
    def add_arguments(cls, parser: argparse.ArgumentParser) -> None:
        for f in fields(cls):
            flag = "--" + f.name.replace("_", "-")
            default = f.default
            typ = type(default) if default is not None else None
            if f.name in ("sample_k",):
                typ = int
            if f.name in ("round_timeout",):
                typ = float
            if f.name in ("checkpoint_dir",):
                typ = str
            parser.add_argument(flag, dest=f.name, type=typ, default=default)

    @classmethod
    def from_args(cls, ns: argparse.Namespace) -> "FederationConfig":
        return cls(**{f.name: getattr(ns, f.name) for f in fields(cls) if hasattr(ns, f.name)})
