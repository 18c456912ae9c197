"""Small pure helpers: key minting, JSON scrubbing, clock seam.

Parity targets (reference file:line):
  * ``random_key`` -- utils.py:38-39 (letters-only tokens, default length 32)
  * ``json_clean`` -- utils.py:23-35 (drop secrets/tensors, stringify datetimes,
    sets -> tuples, recurse into dicts)
"""
from __future__ import annotations

import secrets
import string
from datetime import datetime, timedelta
from typing import Any, Iterable, Mapping

_ALPHABET = string.ascii_letters

#: keys that never leave the process in an introspection response
SECRET_KEYS = ("key", "state_dict")


def random_key(length: int = 32) -> str:
    """Letters-only token.  Same alphabet and default length as the reference,
    but drawn *with* replacement from the OS CSPRNG, so any length is legal and
    the keyspace is 52**length instead of 52!/(52-length)!."""
    if length < 0:
        raise ValueError("length must be non-negative")
    return "".join(secrets.choice(_ALPHABET) for _ in range(length))


def json_clean(data: Any, drop: Iterable[str] = SECRET_KEYS) -> Any:
    """Return a JSON-serialisable copy of ``data`` with secret/tensor entries
    removed.  Unlike the reference this also walks lists/tuples and converts
    tensors/ndarrays that slip through into shapes rather than crashing the
    JSON encoder."""
    drop = tuple(drop)
    if isinstance(data, Mapping):
        out = {}
        for k, v in data.items():
            if k in drop:
                continue
            out[k if isinstance(k, (str, int, float, bool)) or k is None else str(k)] = json_clean(v, drop)
        return out
    if isinstance(data, (set, frozenset)):
        return tuple(json_clean(v, drop) for v in sorted(data, key=str))
    if isinstance(data, (list, tuple)):
        return type(data)(json_clean(v, drop) for v in data) if isinstance(data, list) else tuple(
            json_clean(v, drop) for v in data)
    if isinstance(data, datetime):
        return str(data)
    if isinstance(data, timedelta):
        return data.total_seconds()
    if hasattr(data, "shape") and hasattr(data, "dtype"):
        return {"tensor": list(data.shape), "dtype": str(data.dtype)}
    return data


class Clock:
    """Wall-clock seam.  The reference calls ``datetime.now()`` directly
    (client_manager.py:105,126,130), which makes TTL culling untestable without
    sleeping; every time-dependent component here takes a ``Clock``."""

    def now(self) -> datetime:
        return datetime.now()


class FakeClock(Clock):
    """Manually advanced clock for liveness / timeout tests."""

    def __init__(self, start: datetime | None = None):
        self._now = start or datetime(2026, 1, 1, 0, 0, 0)

    def now(self) -> datetime:
        return self._now

    def advance(self, seconds: float) -> datetime:
        self._now = self._now + timedelta(seconds=seconds)
        return self._now


SYSTEM_CLOCK = Clock()
