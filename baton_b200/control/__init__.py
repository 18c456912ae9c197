"""HTTP control plane (layers L2-L3 of the reference)."""
from .client_manager import ClientManager
from .manager import DEFAULT_N_EPOCH, Experiment, Manager
from .update_manager import (UpdateException, UpdateInProgress, UpdateManager,
                             UpdateNotInProgress)
from .worker import ExperimentWorker

__all__ = [
    "ClientManager", "Experiment", "Manager", "DEFAULT_N_EPOCH", "UpdateManager",
    "UpdateException", "UpdateInProgress", "UpdateNotInProgress", "ExperimentWorker",
]
