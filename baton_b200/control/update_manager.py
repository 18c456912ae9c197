"""Per-round state machine of the parameter server.

Parity target: ``UpdateManager`` and its exception family (reference
update_manager.py:5-68).  Public surface kept: ``start_update(**meta)``,
``end_update()``, ``client_start(id)``, ``client_end(id, resp)``, properties
``in_progress`` / ``clients_left``, ``__len__``, attributes ``update_name``,
``loss_history``, ``n_updates``, ``update_meta``, ``clients``,
``client_responses`` and the name format ``update_{name}_{n:05d}``
(update_manager.py:26).

States: IDLE --start_update--> OPEN --end_update--> IDLE.  The asyncio lock is
still the in-progress flag (update_manager.py:19,31-33) so code that inspects
``update_manager.lock`` keeps working.

Additions: ``client_drop`` (culling a dead participant decrements
``clients_left`` -- quirk 14), ``state()`` for the ``/end_round`` response,
``snapshot``/``restore`` for checkpoint/resume, and per-round timestamps.
"""
from __future__ import annotations

import logging
import time
from asyncio import Lock
from typing import Any, Dict, Optional, Set

from ..utils.misc import random_key

log = logging.getLogger("baton_b200.update")


class UpdateException(Exception):
    """Base class of round state errors."""


class UpdateInProgress(UpdateException):
    """A round is already open."""


class UpdateNotInProgress(UpdateException):
    """No round is open."""


class UpdateManager:
    def __init__(self, name: Optional[str] = None):
        self.lock = Lock()
        self.name = name or random_key(6)
        self.loss_history: list = []
        self.n_updates = 0
        self.round_times: list = []      # wall seconds per finished round
        self._t_open: Optional[float] = None
        self._reset_state()

    # -- bookkeeping -------------------------------------------------------
    def _reset_state(self) -> None:
        self.update_name = "update_{}_{:05d}".format(self.name, self.n_updates)
        self.clients: Set[str] = set()
        self.client_responses: Dict[str, Any] = dict()
        self.update_meta: Optional[dict] = None

    @property
    def in_progress(self) -> bool:
        return self.lock.locked()

    @property
    def clients_left(self) -> int:
        return len(self.clients) - len(self.client_responses)

    def __len__(self) -> int:
        return int(self.in_progress) * len(self.clients)

    # -- transitions -------------------------------------------------------
    async def start_update(self, **update_meta) -> str:
        if self.in_progress:
            raise UpdateInProgress(self.update_name)
        self._reset_state()
        await self.lock.acquire()
        self.update_meta = dict(update_meta)
        self._t_open = time.perf_counter()
        log.info("round open: %s meta=%s", self.update_name, self.update_meta)
        return self.update_name

    def end_update(self) -> Dict[str, Any]:
        if not self.in_progress:
            raise UpdateNotInProgress(self.update_name)
        self.lock.release()
        self.n_updates += 1
        if self._t_open is not None:
            self.round_times.append(time.perf_counter() - self._t_open)
            self._t_open = None
        return self.client_responses

    def client_start(self, client_id: str) -> None:
        if not self.in_progress:
            raise UpdateNotInProgress(client_id)
        self.clients.add(client_id)

    def client_end(self, client_id: str, response: Any) -> None:
        if not self.in_progress:
            raise UpdateNotInProgress(client_id)
        self.client_responses[client_id] = response
        log.info("update received: %s [%d/%d]", client_id,
                 len(self.client_responses), len(self.clients))

    def client_drop(self, client_id: str) -> bool:
        """Remove a participant that died before reporting.  Returns True when
        the participant was pending (so the caller should re-check
        ``clients_left``)."""
        if not self.in_progress or client_id not in self.clients:
            return False
        if client_id in self.client_responses:
            return False
        self.clients.discard(client_id)
        return True

    # -- introspection / persistence ---------------------------------------
    def state(self) -> dict:
        """JSON-friendly round state (what ``/end_round`` was meant to return,
        manager.py:66-68)."""
        return {
            "update_name": self.update_name,
            "in_progress": self.in_progress,
            "clients": set(self.clients),
            "responded": set(self.client_responses),
            "clients_left": self.clients_left,
            "n_updates": self.n_updates,
            "update_meta": dict(self.update_meta or {}),
        }

    def snapshot(self) -> dict:
        return {"name": self.name, "n_updates": self.n_updates,
                "loss_history": list(self.loss_history),
                "update_name": self.update_name}

    def restore(self, snap: dict) -> None:
        if self.in_progress:
            raise UpdateInProgress("cannot restore while a round is open")
        self.n_updates = int(snap.get("n_updates", 0))
        self.loss_history = list(snap.get("loss_history", []))
        self._reset_state()
