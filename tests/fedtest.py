"""In-process federation harness: a manager and N workers, each behind its own
aiohttp TestServer on 127.0.0.1 (the reference tests multi-node the same way --
several processes on one host with distinct ports)."""
import asyncio
import random

import torch
from aiohttp import web
from aiohttp.test_utils import TestClient, TestServer

from baton_b200.control import ExperimentWorker, Manager
from baton_b200.data import linear_regression_shard
from baton_b200.models import LinearModel


def free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class ShardWorker(ExperimentWorker):
    def __init__(self, *a, n=None, seed=0, **kw):
        super().__init__(*a, **kw)
        self.n = n
        self.rng = random.Random(seed)
        self.gen = torch.Generator().manual_seed(seed)
        self.rounds_run = 0

    def get_data(self):
        self.rounds_run += 1
        return linear_regression_shard(self.n, rng=self.rng, generator=self.gen)


class Federation:
    def __init__(self):
        self.servers = []
        self.manager = None
        self.experiment = None
        self.manager_server = None
        self.client = None
        self.workers = []

    async def start_manager(self, model=None, **kw):
        app = web.Application(client_max_size=1 << 30)
        self.manager = Manager(app)
        self.experiment = self.manager.register_experiment(model or LinearModel(), **kw)
        self.manager_server = TestServer(app, host="127.0.0.1")
        await self.manager_server.start_server()
        self.client = TestClient(self.manager_server)
        self.servers.append(self.manager_server)
        return self.experiment

    @property
    def manager_addr(self):
        return "127.0.0.1:{}".format(self.manager_server.port)

    async def add_worker(self, model=None, cls=ShardWorker, wait=True, **kw):
        app = web.Application(client_max_size=1 << 30)
        port = free_port()
        worker = cls(app, model or LinearModel(), self.manager_addr, port=port,
                     heartbeat_time=kw.pop("heartbeat_time", 60), auto_register=False, **kw)
        server = TestServer(app, host="127.0.0.1", port=port)
        await server.start_server()
        self.servers.append(server)
        self.workers.append(worker)
        worker._server = server
        if wait:
            await worker.register_with_manager()
        return worker

    async def wait_round_closed(self, timeout=30.0):
        t0 = asyncio.get_running_loop().time()
        while self.experiment.update_manager.in_progress:
            if asyncio.get_running_loop().time() - t0 > timeout:
                raise TimeoutError("round did not close")
            await asyncio.sleep(0.01)

    async def get(self, path):
        async with self.client.get("/{}/{}".format(self.experiment.name, path)) as resp:
            try:
                body = await resp.json()
            except Exception:
                body = await resp.read()
            return resp.status, body

    async def close(self):
        if self.client is not None:
            await self.client.close()
        for s in self.servers:
            await s.close()
