"""Protocol-compat tier: every row of SURVEY.md section 2.2 against in-process
manager + workers on ephemeral ports."""
import asyncio
import pickle
from collections import OrderedDict

import pytest
import torch

from baton_b200.data import LINEAR_TRUTH
from baton_b200.models import LinearModel, MLP2
from baton_b200.parallel import wire
from conftest import run_async
from fedtest import Federation, ShardWorker


@run_async
async def test_full_round_matches_weighted_mean():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        w1 = await fed.add_worker(n=5, seed=1)
        w2 = await fed.add_worker(n=20, seed=2)
        status, body = await fed.get("start_round?n_epoch=2")
        assert status == 200 and body == {w1.client_id: True, w2.client_id: True}
        await fed.wait_round_closed()
        # global == sum n_k theta_k / N over every state_dict entry
        sd1, sd2 = w1.model.state_dict(), w2.model.state_dict()
        for k, v in exp.model.state_dict().items():
            want = (sd1[k] * 160 + sd2[k] * 640) / 800
            assert torch.allclose(v, want, atol=1e-6), k
        um = exp.update_manager
        assert um.n_updates == 1 and len(um.loss_history) == 2
        want_loss = [(a * 160 + b * 640) / 800 for a, b in zip(w1.last_loss_history, w2.last_loss_history)]
        assert um.loss_history == pytest.approx(want_loss)
        assert exp.client_manager[w1.client_id]["num_updates"] == 1
        assert exp.client_manager[w1.client_id]["last_update"] == "update_lineartest_00000"
        assert w1.n_updates == 1 and not w1.update_in_progress
        status, hist = await fed.get("loss_history")
        assert status == 200 and hist == pytest.approx(want_loss)
        status, st = await fed.get("state")
        assert status == 200 and st["n_updates"] == 1 and st["in_progress"] is False
        status, met = await fed.get("metrics")
        assert status == 200 and met["rounds"] == 1 and met["last"]["n_samples"] == 800
    finally:
        await fed.close()


@run_async
async def test_convergence_over_rounds():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        for s in range(3):
            await fed.add_worker(seed=s, train_kwargs={"lr": 0.02})
        for _ in range(6):
            status, _ = await fed.get("start_round?n_epoch=4")
            assert status == 200
            await fed.wait_round_closed()
        hist = exp.update_manager.loss_history
        assert len(hist) == 24 and hist[-1] < hist[0] * 0.05
        w = exp.model.fc1.weight.detach().flatten()
        assert torch.allclose(w, torch.tensor(LINEAR_TRUTH), atol=0.5)
    finally:
        await fed.close()


@run_async
async def test_start_round_status_codes_and_no_client_lock_release():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        status, body = await fed.get("start_round?n_epoch=abc")
        assert status == 400 and body == {"err": "Invalid Epoch Value"}
        status, body = await fed.get("start_round")
        assert status == 200 and body == []
        assert not exp.update_manager.in_progress      # reference leaks the lock here
        status, body = await fed.get("start_round?n_epoch=1")
        assert status == 200 and body == []

        class Slow(ShardWorker):
            gate = None
            def get_data(self):
                import time
                while not Slow.gate:
                    time.sleep(0.01)
                return super().get_data()
        w = await fed.add_worker(cls=Slow, n=5)
        status, body = await fed.get("start_round")     # default n_epoch = 32
        assert status == 200 and body == {w.client_id: True}
        assert exp.update_manager.update_meta["n_epoch"] == 32
        status, body = await fed.get("start_round?n_epoch=1")
        assert status == 423 and body == {"err": "Update already in progress"}
        # worker is busy -> 409 from its round_start
        import aiohttp
        async with aiohttp.ClientSession() as s:
            url = "http://127.0.0.1:{}/lineartest/round_start?client_id={}&key={}".format(
                w.port, w.client_id, w.key)
            async with s.post(url, data=b"x") as r:
                assert r.status == 409 and (await r.json()) == {"err": "Update in Progress"}
        Slow.gate = True
        await fed.wait_round_closed()
        assert exp.update_manager.n_updates == 3
    finally:
        await fed.close()


@run_async
async def test_update_auth_and_stale_round():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        w = await fed.add_worker(n=5)
        sd = OrderedDict((k, v.clone()) for k, v in w.model.state_dict().items())
        body = pickle.dumps({"state_dict": sd, "n_samples": 10, "update_name": "update_lineartest_00000",
                             "loss_history": [1.0]})
        async with fed.client.post("/lineartest/update", params={"client_id": w.client_id, "key": "bad"},
                                   data=body) as r:
            assert r.status == 401
        async with fed.client.post("/lineartest/update", params={"client_id": "ghost", "key": w.key},
                                   data=body) as r:
            assert r.status == 401
        # authenticated but no round open -> 410 Wrong Update
        async with fed.client.post("/lineartest/update", params={"client_id": w.client_id, "key": w.key},
                                   data=body) as r:
            assert r.status == 410 and (await r.json()) == {"error": "Wrong Update"}
        assert await w.report_update("update_lineartest_00042", 10, [1.0]) == 410
        async with fed.client.post("/lineartest/update", params={"client_id": w.client_id, "key": w.key},
                                   data=b"\x80garbage") as r:
            assert r.status == 400
    finally:
        await fed.close()


@run_async
async def test_worker_round_start_credential_mismatch_evicts_and_reregisters():
    fed = Federation()
    exp = await fed.start_manager()
    try:
        w = await fed.add_worker(n=5)
        old_id = w.client_id
        # manager keeps a record whose key the worker no longer honours
        w.key = "rotated"
        status, body = await fed.get("start_round?n_epoch=1")
        assert status == 200 and body == {old_id: False}
        assert old_id not in exp.client_manager            # 404 -> evicted
        await fed.wait_round_closed()                      # nobody accepted -> round ended
        for _ in range(200):
            if w.client_id != old_id:
                break
            await asyncio.sleep(0.01)
        assert w.client_id != old_id and w.client_id in exp.client_manager
    finally:
        await fed.close()


@run_async
async def test_stock_reference_pickle_payload_roundtrip():
    """A stock ``pickle.dumps`` of the reference schema loads through the safe
    unpickler, and hostile pickles do not."""
    sd = OrderedDict(LinearModel().state_dict())
    blob = pickle.dumps({"state_dict": sd, "update_name": "u", "n_epoch": 3})
    back = wire.loads(blob)
    assert back["n_epoch"] == 3 and torch.equal(back["state_dict"]["fc1.weight"], sd["fc1.weight"])

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("echo pwned",))
    with pytest.raises(pickle.UnpicklingError):
        wire.loads(pickle.dumps({"x": Evil()}))
    assert wire.loads(wire.dumps({"a": 1}, prefer_json=True)) == {"a": 1}
    assert wire.loads(pickle.dumps({"x": Evil()}).replace(b"pwned", b"trust"), trusted=True)


@run_async
async def test_multiple_experiments_share_one_app():
    from aiohttp import web
    from aiohttp.test_utils import TestClient, TestServer
    from baton_b200.control import Manager
    app = web.Application()
    m = Manager(app)
    e1 = m.register_experiment(LinearModel())
    e2 = m.register_experiment(MLP2())
    e3 = m.register_experiment(LinearModel(), name="other")
    with pytest.raises(ValueError):
        m.register_experiment(LinearModel())
    assert [e.name for e in m.experiments] == ["lineartest", "mlp2", "other"] and m["mlp2"] is e2
    server = TestServer(app, host="127.0.0.1"); await server.start_server(); c = TestClient(server)
    try:
        for name in ("lineartest", "mlp2", "other"):
            async with c.get("/{}/clients".format(name)) as r:
                assert r.status == 200 and await r.json() == []
    finally:
        await c.close(); await server.close()


def test_name_falls_back_to_signature_hash():
    from aiohttp import web
    from baton_b200.control import Manager
    class Anon(LinearModel):
        name = None
    m = Manager(web.Application())
    e = m.register_experiment(Anon())
    assert e.name == str(hash(Anon()))


@run_async
async def test_sampling_round_only_notifies_k_clients():
    fed = Federation()
    exp = await fed.start_manager(sample_k=2, seed=5)
    try:
        ws = [await fed.add_worker(n=5, seed=i) for i in range(5)]
        status, body = await fed.get("start_round?n_epoch=1")
        assert status == 200 and len(body) == 2 and all(body.values())
        await fed.wait_round_closed()
        assert sum(w.rounds_run for w in ws) == 2
        assert exp.metrics.records[-1]["n_clients"] == 2
        status, body = await fed.get("start_round?n_epoch=1&sample_k=4")
        assert len(body) == 4
        await fed.wait_round_closed()
    finally:
        await fed.close()
