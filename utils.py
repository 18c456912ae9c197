"""Import shim: ``from utils import PeriodicTask, json_clean, ...`` (reference module name)."""
from baton_b200.utils import (EpochProgress, PeriodicTask, ensure_no_collision,  # noqa: F401
                              json_clean, random_key)
