"""Import shim: ``from manager import Manager, Experiment`` (reference module name)."""
from baton_b200.control.manager import DEFAULT_N_EPOCH, Experiment, Manager  # noqa: F401
