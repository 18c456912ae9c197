"""Runtime helpers (layer L1 of the reference layer map)."""
from .aio import PeriodicTask, ensure_no_collision, run_blocking
from .misc import SYSTEM_CLOCK, Clock, FakeClock, json_clean, random_key
from .progress import EpochProgress

__all__ = [
    "PeriodicTask", "ensure_no_collision", "run_blocking",
    "Clock", "FakeClock", "SYSTEM_CLOCK", "json_clean", "random_key",
    "EpochProgress",
]
