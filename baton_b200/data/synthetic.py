"""Synthetic federated shards.

Parity target: demo ``LinearTestWorker.get_data`` (reference demo.py:52-59):
every round draws ``n in [5, 20]``, ``X ~ N(0,1)`` of shape ``[32n, 10]`` and
``y = (p * X).sum(1)`` for a fixed ground-truth ``p`` (demo.py:55); the varying
``n_samples = 32n`` is what exercises the FedAvg weighting.

Added for the BASELINE.json configs: class-conditional image shards (32x32) and
token shards with three label partitions across K clients -- IID, label-skew
(each client sees ``classes_per_client`` classes) and Dirichlet(alpha) non-IID
(the standard FL benchmark partition; alpha=0.1 is highly skewed).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Sequence, List, Optional, Tuple

import torch

#: ground-truth weights of the reference regression task (demo.py:55)
LINEAR_TRUTH = (11.0, 5.0, 3.0, 2.0, 5.0, 6.0, 2.0, 7.0, 8.0, 1.0)


def linear_regression_shard(n: Optional[int] = None, rng: Optional[random.Random] = None,
                            generator: Optional[torch.Generator] = None,
                            device="cpu") -> Tuple[Tuple[torch.Tensor, torch.Tensor], int]:
    """Fresh regression shard; ``y`` has shape ``(N, 1)`` (fixes quirk 13)."""
    rng = rng or random
    if n is None:
        n = rng.randint(5, 20)
    p = torch.tensor(LINEAR_TRUTH)
    X = torch.randn(32 * n, len(LINEAR_TRUTH), generator=generator)
    y = (p * X).sum(1, keepdim=True)
    return (X.to(device), y.to(device)), 32 * n


@dataclass
class ShardSpec:
    """Label distribution of one client."""
    client: int
    class_probs: torch.Tensor  # [num_classes], sums to 1
    n_samples: int


def iid_label_shards(k: int, num_classes: int, n_samples: "int | Sequence[int]") -> List[ShardSpec]:
    ns = [n_samples] * k if isinstance(n_samples, int) else list(n_samples)
    p = torch.full((num_classes,), 1.0 / num_classes)
    return [ShardSpec(i, p.clone(), ns[i]) for i in range(k)]


def label_skew_shards(k: int, num_classes: int, n_samples: "int | Sequence[int]",
                      classes_per_client: int = 2) -> List[ShardSpec]:
    ns = [n_samples] * k if isinstance(n_samples, int) else list(n_samples)
    out = []
    for i in range(k):
        p = torch.zeros(num_classes)
        for j in range(classes_per_client):
            p[(i * classes_per_client + j) % num_classes] = 1.0 / classes_per_client
        out.append(ShardSpec(i, p, ns[i]))
    return out


def dirichlet_label_shards(k: int, num_classes: int, n_samples: "int | Sequence[int]",
                           alpha: float = 0.1, seed: int = 0) -> List[ShardSpec]:
    """Per-client class proportions ~ Dirichlet(alpha * 1)."""
    ns = [n_samples] * k if isinstance(n_samples, int) else list(n_samples)
    g = torch.Generator().manual_seed(seed)
    conc = torch.full((num_classes,), float(alpha))
    # torch.distributions has no generator argument; sample gammas directly
    gam = torch._standard_gamma(conc.expand(k, num_classes).contiguous(), generator=g).clamp_min(1e-30)
    probs = gam / gam.sum(1, keepdim=True)
    return [ShardSpec(i, probs[i], ns[i]) for i in range(k)]


def _labels_for(spec: ShardSpec, generator: Optional[torch.Generator]) -> torch.Tensor:
    return torch.multinomial(spec.class_probs, spec.n_samples, replacement=True, generator=generator)


def image_shard(spec: ShardSpec, *, channels: int = 3, size: int = 32, seed: int = 0,
                dtype=torch.float32, channels_last: bool = True, pin: bool = False,
                noise: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Class-conditional Gaussian images: each class has a fixed random mean
    pattern (shared across clients through ``seed``), samples are mean + noise.
    Returned as NHWC when ``channels_last`` (the layout the conv kernels use)."""
    num_classes = spec.class_probs.numel()
    gm = torch.Generator().manual_seed(seed)
    means = torch.randn(num_classes, size, size, channels, generator=gm) * 0.5
    g = torch.Generator().manual_seed(seed * 7919 + 1000003 * (spec.client + 1))
    y = _labels_for(spec, g)
    X = means[y] + noise * torch.randn(spec.n_samples, size, size, channels, generator=g)
    if not channels_last:
        X = X.permute(0, 3, 1, 2).contiguous()
    X = X.to(dtype)
    if pin and torch.cuda.is_available():
        X, y = X.pin_memory(), y.pin_memory()
    return X, y


def token_shard(spec: ShardSpec, *, seq_len: int = 128, vocab: int = 30522, seed: int = 0,
                pin: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Class-conditional token sequences for the BERT config: each class
    over-samples a class-specific slice of the vocabulary."""
    num_classes = spec.class_probs.numel()
    g = torch.Generator().manual_seed(seed * 104729 + 1000003 * (spec.client + 1))
    y = _labels_for(spec, g)
    base = torch.randint(0, vocab, (spec.n_samples, seq_len), generator=g)
    band = max(1, vocab // (num_classes * 4))
    hot = torch.randint(0, band, (spec.n_samples, seq_len), generator=g) + (y.view(-1, 1) * band)
    use_hot = torch.rand(spec.n_samples, seq_len, generator=g) < 0.3
    X = torch.where(use_hot, hot, base)
    if pin and torch.cuda.is_available():
        X, y = X.pin_memory(), y.pin_memory()
    return X, y
