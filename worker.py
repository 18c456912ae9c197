"""Import shim: ``from worker import ExperimentWorker`` (reference module name)."""
from baton_b200.control.worker import ExperimentWorker  # noqa: F401
