"""Fault-injection tier: dead workers, partial
aggregation, round timeout, manager restart, checkpoint/resume."""
import asyncio

import pytest
import torch

from baton_b200 import ckpt
from baton_b200.models import LinearModel
from baton_b200.utils import FakeClock
from conftest import run_async
from fedtest import Federation


@run_async
async def test_straggler_cut_off_by_end_round_partial_aggregation():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        good = await fed.add_worker(n=5, seed=1)
        dead = await fed.add_worker(n=20, seed=2)
        dead.drop_next_reports = 1                  # trains but never reports
        status, body = await fed.get("start_round?n_epoch=1")
        assert all(body.values())
        for _ in range(500):
            if len(exp.update_manager.client_responses) == 1 and not dead.update_in_progress:
                break
            await asyncio.sleep(0.01)
        assert exp.update_manager.in_progress and exp.update_manager.clients_left == 1
        status, st = await fed.get("end_round")     # operator straggler cut-off (500 in the reference)
        assert status == 200 and st["in_progress"] is False and st["n_updates"] == 1
        for k, v in exp.model.state_dict().items():
            assert torch.allclose(v, good.model.state_dict()[k])   # only the reporter counts
        assert len(exp.update_manager.loss_history) == 1
        # the straggler's late report is rejected as stale
        assert await dead.report_update("update_lineartest_00000", 640, [0.1]) == 410
        status, st = await fed.get("end_round")     # idempotent when idle
        assert status == 200
    finally:
        await fed.close()


@run_async
async def test_round_timeout_auto_ends_round():
    fed = Federation()
    exp = await fed.start_manager(round_timeout=0.3)
    try:
        a = await fed.add_worker(n=5, seed=1)
        b = await fed.add_worker(n=5, seed=2)
        b.drop_next_reports = 1
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed(timeout=5)
        assert exp.update_manager.n_updates == 1
        assert exp.metrics.records[-1]["n_clients"] == 1 and exp.metrics.records[-1]["participants"] == 2
    finally:
        await fed.close()


@run_async
async def test_training_failure_and_eviction_unblock_round():
    clock = FakeClock()
    fed = Federation()
    exp = await fed.start_manager(clock=clock, client_ttl=300)
    try:
        a = await fed.add_worker(n=5, seed=1)
        b = await fed.add_worker(n=5, seed=2)
        b.fail_next_rounds = 1                      # raises inside local training
        await fed.get("start_round?n_epoch=1")
        for _ in range(500):
            if len(exp.update_manager.client_responses) == 1 and not b.update_in_progress:
                break
            await asyncio.sleep(0.01)
        assert exp.update_manager.in_progress
        # b stops heart-beating; a keeps going; TTL cull drops b from the participant set
        clock.advance(200)
        await a.heartbeat()
        clock.advance(200)
        await exp.client_manager.cull_clients()
        await fed.wait_round_closed(timeout=5)
        assert b.client_id not in exp.client_manager and a.client_id in exp.client_manager
        assert exp.update_manager.n_updates == 1
        assert not b.update_in_progress             # busy flag cleared even on failure
    finally:
        await fed.close()


@run_async
async def test_manager_restart_worker_reregisters_and_resume_from_checkpoint(tmp_path):
    fed = Federation()
    exp = await fed.start_manager(checkpoint_dir=str(tmp_path))
    try:
        w = await fed.add_worker(n=5, seed=3)
        await fed.get("start_round?n_epoch=2")
        await fed.wait_round_closed()
        for _ in range(500):                      # the checkpoint is written right AFTER the round lock is released
            if exp.last_checkpoint:
                break
            await asyncio.sleep(0.01)
        path = exp.last_checkpoint
        assert path and path.endswith("lineartest_00001.pt")
        payload = torch.load(path, weights_only=True)
        assert set(payload) >= {"state_dict", "n_updates", "loss_history", "update_name"}
        stock = torch.nn.Linear(10, 1)              # plain-PyTorch layout stays loadable
        stock.load_state_dict({k.replace("fc1.", ""): v for k, v in payload["state_dict"].items()})
        saved_weights = {k: v.clone() for k, v in exp.model.state_dict().items()}
        # "restart": the manager forgets every client -> heartbeat gets 401 -> re-register
        old = w.client_id
        exp.client_manager.clients.clear()
        assert await w.heartbeat() is False
        for _ in range(300):
            if w.client_id != old:
                break
            await asyncio.sleep(0.01)
        assert w.client_id != old and w.client_id in exp.client_manager
    finally:
        await fed.close()
    # a new manager process resumes the counter, the loss history and the weights
    fed2 = Federation()
    exp2 = await fed2.start_manager(checkpoint_dir=str(tmp_path), resume=True)
    try:
        assert exp2.update_manager.n_updates == 1
        assert exp2.update_manager.update_name == "update_lineartest_00001"
        assert len(exp2.update_manager.loss_history) == 2
        for k, v in exp2.model.state_dict().items():
            assert torch.equal(v, saved_weights[k])
    finally:
        await fed2.close()


def test_checkpoint_rotation(tmp_path):
    from baton_b200.control import UpdateManager
    um = UpdateManager("m")
    model = LinearModel()
    for i in range(5):
        um.n_updates = i + 1
        ckpt.save_checkpoint(str(tmp_path), "m", model, um, keep=2)
    names = [p.rsplit("/", 1)[-1] for p in ckpt.list_checkpoints(str(tmp_path), "m")]
    assert names == ["m_00004.pt", "m_00005.pt"]
    assert ckpt.latest_checkpoint(str(tmp_path), "m").endswith("m_00005.pt")
    assert ckpt.latest_checkpoint(str(tmp_path), "zzz") is None


@run_async
async def test_heartbeat_backoff_then_recover():
    """Manager unreachable -> the worker backs off and recovers when it returns."""
    from baton_b200.control import ExperimentWorker
    from aiohttp import web
    import baton_b200.control.worker as wmod
    sleeps = []
    real_sleep = asyncio.sleep

    async def fake_sleep(t):
        sleeps.append(t)
        await real_sleep(0)
    fed = Federation()
    exp = await fed.start_manager()
    w = await fed.add_worker(n=5)
    good_url = w.manager_url
    try:
        w.manager_url = "http://127.0.0.1:1/lineartest/"
        wmod.asyncio.sleep = fake_sleep
        task = asyncio.ensure_future(w.heartbeat())
        for _ in range(200):
            if len(sleeps) >= 4:
                break
            await real_sleep(0.005)
        w.manager_url = good_url
        assert await asyncio.wait_for(task, 5) is True
        assert sleeps[:4] == [1.0, 2.0, 4.0, 8.0]   # exponential, starts at 1 s (worker.py:59,76-79)
    finally:
        wmod.asyncio.sleep = real_sleep
        await fed.close()


@run_async
async def test_round_does_not_close_while_the_fan_out_is_still_in_flight():
    """A fast client trains and reports while a slow peer has not even answered its round_start yet (e.g. it is being
    sent the whole model): the round must wait for the fan-out to finish instead of closing on "everybody registered
    so far has reported" -- found on 2 GPUs, where the fast seat finished 4 SGD steps before the 45 MB POST to the
    returning seat completed (tests/mp_api_check.py)."""
    import asyncio

    fed = Federation()
    exp = await fed.start_manager()
    try:
        fast = await fed.add_worker(n=5, seed=0)
        slow = await fed.add_worker(n=5, seed=1)
        orig = slow.round_start

        async def delayed(request):
            for _ in range(2000):               # the slow seat answers its round_start only AFTER the fast one reported
                if exp.update_manager.client_responses:
                    break
                await asyncio.sleep(0.01)
            await asyncio.sleep(0.05)
            return await orig(request)
        # re-route the slow worker's handler
        for r in slow.app.router.routes():
            if r.resource.canonical.endswith("/round_start"):
                r._handler = delayed
        status, body = await fed.get("start_round?n_epoch=1")
        assert status == 200 and all(body.values()) and len(body) == 2
        await fed.wait_round_closed()
        rec = exp.metrics.records[-1]
        assert rec["n_clients"] == 2, rec           # both updates were aggregated, not just the fast one
    finally:
        await fed.close()
