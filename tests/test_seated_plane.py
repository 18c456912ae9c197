"""The metadata-only ("seated") data plane through the real HTTP control plane.

CPU tier: sessions are faked (a shared dict stands in for the symmetric memory) so the protocol --
JSON updates without tensors, the per-rank aggregation plan, alive ranks, the manager's lazy
``pull_global`` and checkpointing -- is exercised without a GPU.  GPU tier: one seat on cuda:0 with
the real fused kernel."""
import asyncio

import pytest
import torch

from baton_b200.models import LinearModel
from baton_b200.parallel import wire
from conftest import run_async
from fedtest import Federation, ShardWorker


class FakeFabric:
    def __init__(self):
        self.seats = {}
        self.calls = []


class FakeSession:
    """Stand-in for FedAvgSession: same ``aggregate`` contract, reduce done on the host."""

    def __init__(self, fabric, rank, model):
        self.fabric, self.rank, self.model, self.device = fabric, rank, model, "cpu"
        fabric.seats[rank] = self
        self._pending = None

    def aggregate(self, n_samples_by_rank, alive_ranks=None):
        self.fabric.calls.append((self.rank, list(n_samples_by_rank), list(alive_ranks or [])))
        alive = list(alive_ranks or range(len(n_samples_by_rank)))
        if self.rank not in alive:
            return
        total = sum(n_samples_by_rank[r] for r in alive)
        # every seat computes the same weighted mean from the (frozen) pre-round replicas
        snap = self.fabric.snapshot
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(sum(snap[r][k] * (n_samples_by_rank[r] / total) for r in alive if n_samples_by_rank[r] > 0))


@run_async
async def test_seated_plane_round_trip_and_pull_global(tmp_path):
    fed = Federation()
    exp = await fed.start_manager(dataplane="fused", checkpoint_dir=str(tmp_path))
    fabric = FakeFabric()
    try:
        workers = []
        for r, n in enumerate((5, 20, 10)):
            m = LinearModel()
            m.load_state_dict(exp.model.state_dict())
            w = await fed.add_worker(model=m, n=n, seed=r, dataplane="fused", session=FakeSession(fabric, r, m))
            workers.append(w)
        assert [exp.client_manager[w.client_id]["rank"] for w in workers] == [0, 1, 2]
        assert exp.client_manager[workers[0].client_id]["backend"] == "fused"

        # freeze the post-training replicas right before the manager fans out the plan
        orig = exp.plane.aggregate

        async def spy(experiment, responses):
            fabric.snapshot = {w.plane.rank: {k: v.clone() for k, v in w.model.state_dict().items()} for w in workers}
            for d in responses.values():
                assert "state_dict" not in d and set(d) >= {"n_samples", "update_name", "loss_history", "rank"}
            return await orig(experiment, responses)
        exp.plane.aggregate = spy
        status, body = await fed.get("start_round?n_epoch=2")
        assert status == 200 and all(body.values())
        await fed.wait_round_closed()
        # every seat ran the collective with the same plan
        plans = sorted(fabric.calls)
        assert [p[0] for p in plans] == [0, 1, 2]
        assert all(p[1] == [160.0, 640.0, 320.0] and p[2] == [0, 1, 2] for p in plans)
        want = {k: sum(fabric.snapshot[r][k] * n for r, n in ((0, 160), (1, 640), (2, 320))) / 1120
                for k in fabric.snapshot[0]}
        for w in workers:
            for k, v in w.model.state_dict().items():
                assert torch.allclose(v, want[k], atol=1e-6)
        # the manager's copy is stale until asked; /state_dict (and the checkpoint) pull it from seat 0
        for _ in range(500):            # the checkpoint is written right after the round lock is released
            if exp.last_checkpoint:
                break
            await asyncio.sleep(0.01)
        assert exp.last_checkpoint and exp.model_is_stale is False          # the checkpoint pulled it
        async with fed.client.get("/lineartest/state_dict") as r:
            sd = wire.loads(await r.read())["state_dict"]
        for k in want:
            assert torch.allclose(sd[k], want[k], atol=1e-6)
            assert torch.allclose(exp.model.state_dict()[k], want[k], atol=1e-6)
        payload = torch.load(exp.last_checkpoint, weights_only=True)
        assert torch.allclose(payload["state_dict"]["fc1.weight"], want["fc1.weight"], atol=1e-6)
        assert len(exp.update_manager.loss_history) == 2
        # bytes over HTTP are metadata-sized, not model-sized
        assert exp.metrics.records[-1]["bytes_http"] < 4096
    finally:
        await fed.close()


@run_async
async def test_seated_plane_excludes_dead_seat_from_plan():
    fed = Federation()
    exp = await fed.start_manager(dataplane="fused")
    fabric = FakeFabric()
    try:
        ws = []
        for r in range(3):
            m = LinearModel()
            ws.append(await fed.add_worker(model=m, n=5, seed=r, dataplane="fused", session=FakeSession(fabric, r, m)))
        ws[2].drop_next_reports = 1                 # seat 2 trains but never reports
        await fed.get("start_round?n_epoch=1")
        for _ in range(500):
            if len(exp.update_manager.client_responses) == 2 and not ws[2].update_in_progress:
                break
            await asyncio.sleep(0.01)
        fabric.snapshot = {w.plane.rank: {k: v.clone() for k, v in w.model.state_dict().items()} for w in ws}
        status, _ = await fed.get("end_round")
        assert status == 200
        plans = sorted(fabric.calls)
        assert [p[0] for p in plans] == [0, 1, 2]                  # the straggler still receives the broadcast
        assert all(p[1] == [160.0, 160.0, 0.0] for p in plans)     # ... but contributes weight 0
        # now the seat dies for real: evicted -> not in alive_ranks any more
        exp.client_manager.evict(ws[2].client_id, "test")
        fabric.calls.clear()
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed()
        assert sorted(p[0] for p in fabric.calls) == [0, 1]
        assert all(p[2] == [0, 1] for p in fabric.calls)
    finally:
        await fed.close()


@pytest.mark.gpu
@run_async
async def test_http_control_plane_drives_fused_kernel_on_gpu():
    """Manager + one GPU seat in-process: HTTP carries metadata, the fused kernel applies the round."""
    from baton_b200.control.gpu_worker import GpuExperimentWorker
    from baton_b200.data import ShardSpec, image_shard
    from baton_b200.models import resnet18
    from fedtest import free_port
    from aiohttp import web
    from aiohttp.test_utils import TestServer
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    fed = Federation()
    exp = await fed.start_manager(model=resnet18(10), dataplane="fused")
    try:
        X, y = image_shard(ShardSpec(0, torch.full((10,), 0.1), 512), dtype=torch.bfloat16, noise=0.3)
        app = web.Application(client_max_size=1 << 30)
        port = free_port()
        model = resnet18(10)
        w = GpuExperimentWorker(app, model, fed.manager_addr, device=dev, shard_fn=lambda: (X, y), port=port,
                                auto_register=False, train_kwargs={"lr": 0.05, "batch_size": 128})
        server = TestServer(app, host="127.0.0.1", port=port)
        await server.start_server()
        fed.servers.append(server)
        await w.register_with_manager()
        g0 = w.arena.global_w.clone()
        for rnd in range(2):
            status, body = await fed.get("start_round?n_epoch=2")
            assert status == 200 and all(body.values())
            await fed.wait_round_closed(timeout=120)
        w.fed_session.check()
        assert w.fed_session.rounds == 2
        assert not torch.equal(w.arena.global_w, g0)                       # the global model moved
        assert torch.equal(w.arena.theta, w.arena.global_w)                # replica == global after the round
        hist = exp.update_manager.loss_history
        assert len(hist) == 4 and hist[-1] < hist[0]
        async with fed.client.get("/resnet18/state_dict") as r:            # manager pulls from the seat
            sd = wire.loads(await r.read())["state_dict"]
        assert torch.allclose(sd["fc.weight"], model.fc.weight.detach().cpu())
        assert exp.metrics.records[-1]["bytes_http"] < 8192
    finally:
        await fed.close()


@run_async
async def test_seated_plane_distributes_the_model_to_new_and_returning_seats():
    """A seat that has never been given the global model (first round; re-registration after an eviction) receives
    the full state_dict inside its round_start -- pulled from a live seat when the manager's copy is stale -- and the
    aggregation plan carries the manager's round index (the seats derive the collective's barrier epoch from it)."""
    fed = Federation()
    exp = await fed.start_manager(dataplane="fused")
    fabric = FakeFabric()
    try:
        ws = []
        for r in range(2):
            torch.manual_seed(100 + r)
            m = LinearModel()                                    # different weights on every seat on purpose
            ws.append(await fed.add_worker(model=m, n=5, seed=r, dataplane="fused", session=FakeSession(fabric, r, m)))
        loaded = {0: [], 1: []}
        for r, w in enumerate(ws):
            orig = w.model.load_state_dict
            w.model.load_state_dict = (lambda sd, _o=orig, _r=r, **kw: (loaded[_r].append({k: v.clone() for k, v in sd.items()}),
                                                                      _o(sd, **kw))[1])
        init = {k: v.clone() for k, v in exp.model.state_dict().items()}
        orig_agg = exp.plane.aggregate
        plans = []

        async def spy(experiment, responses):
            fabric.snapshot = {w.plane.rank: {k: v.clone() for k, v in w.model.state_dict().items()} for w in ws}
            plans.append(exp.plane.n_aggregates)
            return await orig_agg(experiment, responses)
        exp.plane.aggregate = spy
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed()
        assert len(loaded[0]) == 1 and len(loaded[1]) == 1                       # round 1: everybody got the manager's model
        for k in init:
            assert torch.equal(loaded[0][0][k], init[k]) and torch.equal(loaded[1][0][k], init[k])
        for k, v in ws[0].model.state_dict().items():
            assert torch.allclose(v, ws[1].model.state_dict()[k])
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed()
        assert len(loaded[0]) == 1 and len(loaded[1]) == 1                       # round 2: metadata only
        assert plans == [0, 1]
        # seat 1 is evicted, misses a round, comes back under a new client id
        exp.client_manager.evict(ws[1].client_id, "test")
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed()
        current = {k: v.clone() for k, v in ws[0].model.state_dict().items()}     # the global model after round 3
        with torch.no_grad():
            for v in ws[1].model.state_dict().values():
                v.add_(1.0)                                                       # its replica is stale / wrong
        ws[1].client_id = ws[1].key = None
        await ws[1].register_with_manager()
        await fed.get("start_round?n_epoch=1")
        await fed.wait_round_closed()
        assert len(loaded[0]) == 1 and len(loaded[1]) == 2                       # only the returning seat was re-synced
        for k in current:
            assert torch.allclose(loaded[1][1][k], current[k], atol=1e-6), k      # ... with the CURRENT global model
        for k, v in ws[0].model.state_dict().items():
            assert torch.allclose(v, ws[1].model.state_dict()[k], atol=1e-6)
        assert plans == [0, 1, 2, 3]
    finally:
        await fed.close()
