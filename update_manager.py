"""Import shim: ``from update_manager import UpdateManager, UpdateException`` (reference module name)."""
from baton_b200.control.update_manager import (UpdateException, UpdateInProgress,  # noqa: F401
                                                UpdateManager, UpdateNotInProgress)
