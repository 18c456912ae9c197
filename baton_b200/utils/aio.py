"""asyncio helpers for the control plane.

Behavioural parity targets (reference file:line):
  * ``ensure_no_collision``  -- utils.py:11-20  (drop-if-running guard)
  * ``PeriodicTask``         -- utils.py:42-67  (sleep/await loop with start/stop)

Differences by design:
  * the guard is per *instance* (two workers living in one process no longer
    block each other), and
  * a ``PeriodicTask`` may be constructed before an event loop exists; the
    coroutine is only scheduled by ``start()``, which the owners call from
    aiohttp ``on_startup`` hooks.
"""
from __future__ import annotations

import asyncio
import functools
import logging
import weakref
from contextlib import suppress
from typing import Any, Awaitable, Callable, Optional

log = logging.getLogger("baton_b200.aio")


class _Guard:
    """One in-flight marker per (owner, function)."""

    __slots__ = ("busy", "__weakref__")

    def __init__(self) -> None:
        self.busy = False


def ensure_no_collision(fxn: Callable[..., Awaitable[Any]]):
    """Make an async callable non-reentrant: a second call that arrives while
    the first is still running returns ``None`` immediately instead of queueing.

    For bound methods the guard is keyed on ``self`` so separate instances do
    not collide; plain coroutine functions share one module-level guard.
    """
    per_owner: "weakref.WeakKeyDictionary[Any, _Guard]" = weakref.WeakKeyDictionary()
    free_guard = _Guard()

    def _guard_for(args) -> _Guard:
        if args:
            owner = args[0]
            try:
                guard = per_owner.get(owner)
                if guard is None:
                    guard = per_owner[owner] = _Guard()
                return guard
            except TypeError:  # unhashable / not weak-referenceable first arg
                pass
        return free_guard

    @functools.wraps(fxn)
    async def guarded(*args, **kwargs):
        guard = _guard_for(args)
        if guard.busy:
            log.debug("%s already running; call dropped", fxn.__name__)
            return None
        guard.busy = True
        try:
            return await fxn(*args, **kwargs)
        finally:
            guard.busy = False

    guarded.__wrapped_guard__ = _guard_for  # introspection hook for tests
    return guarded


class PeriodicTask:
    """Call ``func`` every ``time`` seconds until stopped.

    ``start()`` returns ``self`` so the reference idiom
    ``PeriodicTask(f, t).start()`` keeps working (client_manager.py:23-24).
    Exceptions raised by ``func`` are logged and do not kill the loop: a failed
    heartbeat must not silently end all later heartbeats.
    """

    def __init__(self, func: Callable[[], Awaitable[Any]], time: float,
                 *, run_immediately: bool = False,
                 sleep: Optional[Callable[[float], Awaitable[None]]] = None):
        self.func = func
        self.time = time
        self.is_started = False
        self.run_immediately = run_immediately
        self.n_calls = 0
        self._sleep = sleep or asyncio.sleep
        self._task: Optional[asyncio.Task] = None

    def start(self) -> "PeriodicTask":
        if not self.is_started:
            self.is_started = True
            self._task = asyncio.ensure_future(self._run())
        return self

    async def stop(self) -> None:
        if self.is_started:
            self.is_started = False
            task, self._task = self._task, None
            if task is not None and task is not asyncio.current_task():
                task.cancel()
                with suppress(asyncio.CancelledError):
                    await task

    async def _tick(self) -> None:
        try:
            await self.func()
        except asyncio.CancelledError:
            raise
        except Exception:  # pragma: no cover - defensive; logged for operators
            log.exception("periodic task %r failed", getattr(self.func, "__name__", self.func))
        self.n_calls += 1

    async def _run(self) -> None:
        if self.run_immediately and self.is_started:
            await self._tick()
        while self.is_started:
            await self._sleep(self.time)
            if not self.is_started:
                break
            await self._tick()


async def run_blocking(fn: Callable[..., Any], *args, executor=None, **kwargs):
    """Run a blocking callable (local training) off the event loop so the
    heartbeat and HTTP handlers stay responsive (fixes worker.py:105 which
    blocks the loop for the whole local epoch)."""
    loop = asyncio.get_running_loop()
    return await loop.run_in_executor(executor, functools.partial(fn, *args, **kwargs))
