import asyncio
import string
from datetime import datetime

import pytest
import torch

from baton_b200.utils import (EpochProgress, FakeClock, PeriodicTask, ensure_no_collision,
                              json_clean, random_key)
from conftest import run_async


def test_random_key_alphabet_and_length():
    k = random_key()
    assert len(k) == 32 and set(k) <= set(string.ascii_letters)
    assert len(random_key(6)) == 6
    assert len(random_key(100)) == 100          # the reference caps at 52 (sampling w/o replacement)
    assert random_key() != random_key()


def test_json_clean_drops_secrets_and_converts():
    now = datetime(2026, 1, 2, 3, 4, 5)
    data = {"key": "secret", "client_id": "c", "state_dict": {"w": 1}, "when": now,
            "tags": {"b", "a"}, "nested": {"key": "x", "ok": 1, "deep": {"state_dict": 2, "v": [now]}}}
    out = json_clean(data)
    assert "key" not in out and "state_dict" not in out
    assert out["when"] == str(now)
    assert out["tags"] == ("a", "b")
    assert out["nested"] == {"ok": 1, "deep": {"v": [str(now)]}}
    import json
    json.dumps(out)


def test_json_clean_tensor_is_summarised():
    out = json_clean({"t": torch.zeros(2, 3)})
    assert out["t"]["tensor"] == [2, 3]


def test_epoch_progress_is_a_true_mean():
    ep = EpochProgress(0, range(4), verbose=False)
    for _ in ep:
        ep.update_loss(4.0)
    assert ep.loss == pytest.approx(4.0)        # reference recurrence reports 4.867 here
    ep = EpochProgress(0, range(3), verbose=False)
    for i, _ in enumerate(ep):
        ep.update_loss(torch.tensor(float(i)))
    assert ep.loss == pytest.approx(1.0)
    assert ep.N == 3


def test_epoch_progress_empty():
    ep = EpochProgress(0, [], verbose=False)
    assert list(ep) == [] and ep.loss == 0.0


@run_async
async def test_periodic_task_runs_and_stops():
    calls = []

    async def tick():
        calls.append(1)

    task = PeriodicTask(tick, 0.01).start()
    assert task.is_started
    for _ in range(600):                 # poll instead of a fixed sleep: robust on a loaded machine
        if len(calls) >= 2:
            break
        await asyncio.sleep(0.005)
    await task.stop()
    n = len(calls)
    assert n >= 2
    await asyncio.sleep(0.03)
    assert len(calls) == n and not task.is_started
    await task.stop()  # idempotent


@run_async
async def test_periodic_task_survives_exceptions():
    calls = []

    async def tick():
        calls.append(1)
        raise RuntimeError("boom")

    task = PeriodicTask(tick, 0.01).start()
    for _ in range(600):
        if len(calls) >= 2:
            break
        await asyncio.sleep(0.005)
    await task.stop()
    assert len(calls) >= 2


@run_async
async def test_ensure_no_collision_drops_reentrant_call_per_instance():
    class W:
        def __init__(self):
            self.n = 0
            self.gate = asyncio.Event()

        @ensure_no_collision
        async def work(self):
            self.n += 1
            await self.gate.wait()
            return "done"

    a, b = W(), W()
    ta = asyncio.ensure_future(a.work())
    await asyncio.sleep(0)
    assert await a.work() is None               # second call on the same instance is dropped
    tb = asyncio.ensure_future(b.work())        # other instance is NOT blocked (reference bug)
    await asyncio.sleep(0)
    assert b.n == 1
    a.gate.set(), b.gate.set()
    assert await ta == "done" and await tb == "done"
    a.gate.set()
    assert await a.work() == "done"             # guard released afterwards
    assert a.n == 2


def test_fake_clock():
    c = FakeClock()
    t0 = c.now()
    c.advance(301)
    assert (c.now() - t0).total_seconds() == 301
