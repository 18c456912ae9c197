"""Data plane: aggregation algebra, wire formats, transports and the fused
NVLink FedAvg collective.  CUDA-dependent modules (``arena``, ``symm``,
``fedavg``, ``engine``) are imported lazily so the control plane works on a
GPU-less host."""
from . import wire
from .aggregate import client_weights, fedavg_into, fedavg_loss_history

__all__ = ["wire", "client_weights", "fedavg_into", "fedavg_loss_history"]
