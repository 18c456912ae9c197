"""Import shim: ``from client_manager import ClientManager`` (reference module name)."""
from baton_b200.control.client_manager import ClientManager  # noqa: F401
