"""Federated client: register, heartbeat, receive a round, train locally, report.

Parity target: ``ExperimentWorker`` (reference worker.py:12-127).
  * constructor signature            worker.py:13-32
        ExperimentWorker(app, model, manager, name=None, port=8080,
                         heartbeat_time=60, worker_host=None)
  * ``register_with_manager``        worker.py:40-55   GET {manager}/register {url, port}
  * ``heartbeat``                    worker.py:57-79   200 ok | 401 re-register |
                                                      else exponential backoff from 1 s
  * ``round_start`` handler          worker.py:87-101  POST /{name}/round_start?client_id&key
        409 while busy, 404 on credential mismatch (+ re-register), else load the
        weights, start training in the background and answer 200 "OK" at once
  * ``_run_round``                   worker.py:103-106 get_data -> model.train -> report_update
  * ``report_update``                worker.py:108-124 POST {manager}/update; 401 re-register,
                                                      410 = stale round
  * ``get_data`` (abstract)          worker.py:126-127 -> (args_for_train, n_samples)

Fixed: the busy flag is really set/cleared (quirk 3);
local training runs in an executor so heartbeats and HTTP stay live during the
epoch (quirk 4); backoff is capped; the HTTP session is closed on cleanup
(quirk 15); registration starts from ``on_startup`` rather than from
``__init__`` (quirk 18).

New: data planes (``http`` | ``fused`` | ``nccl``), ``POST /{name}/aggregate``
and ``GET /{name}/state_dict`` for GPU-seated clients, fault-injection hooks.
The user model may expose the training entry point as ``train(*data, n_epoch=)``
(the reference contract, demo.py:29), ``fit`` or ``local_train`` (quirk 11).
"""
from __future__ import annotations

import asyncio
import logging
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Optional
from urllib.parse import urljoin

import aiohttp
from aiohttp import web

from ..parallel import wire
from ..parallel.dataplane import WorkerPlane, make_worker_plane
from ..utils.aio import PeriodicTask, ensure_no_collision, run_blocking

log = logging.getLogger("baton_b200.worker")

MAX_BACKOFF = 60.0


def resolve_train_fn(model) -> Callable:
    """Find the local-training entry point of a user model."""
    for attr in ("local_train", "fit"):
        fn = getattr(model, attr, None)
        if callable(fn):
            return fn
    fn = getattr(model, "train", None)
    import torch.nn as nn
    if fn is not None and getattr(type(model), "train", None) is not nn.Module.train:
        return fn  # the reference contract: train() overridden as the SGD loop
    raise TypeError("model must define local_train/fit, or override train(*data, n_epoch=...)")


class ExperimentWorker:
    def __init__(self, app: web.Application, model, manager: str, name: Optional[str] = None,
                 port: int = 8080, heartbeat_time: float = 60, worker_host: Optional[str] = None, *,
                 dataplane: Any = "http", session=None, auto_register: bool = True,
                 trusted_peers: bool = False, train_kwargs: Optional[dict] = None):
        self.name = str(name or getattr(model, "name", None) or hash(model))
        self.model = model
        self.app = app
        self.port = port
        self.worker_host = worker_host
        self.manager = manager
        self.manager_url = "http://{}/{}/".format(manager, self.name)
        self.plane: WorkerPlane = make_worker_plane(dataplane, session)
        self.trusted_peers = trusted_peers
        self.train_kwargs = dict(train_kwargs or {})
        self.n_updates = 0
        self.update_in_progress = False
        self.last_update: Optional[str] = None
        self.client_id: Optional[str] = None
        self.key: Optional[str] = None
        self.heartbeat_time = heartbeat_time
        self.last_loss_history: list = []
        self._session: Optional[aiohttp.ClientSession] = None
        self._heartbeat_manager: Optional[PeriodicTask] = None
        # one training thread; a GPU-seated worker (``self.device`` set by the subclass before this constructor runs)
        # binds it to its own CUDA device -- new threads start on device 0
        self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="baton-train",
                                            initializer=self._bind_device)
        self._round_task: Optional[asyncio.Task] = None
        self._auto_register = auto_register
        # fault-injection seams used by the test-suite
        self.fail_next_rounds = 0          # raise inside local training
        self.drop_next_reports = 0         # train but never report (straggler / death)
        self.register_handlers()
        app.on_startup.append(self._on_startup)
        app.on_cleanup.append(self._on_cleanup)

    # -- lifecycle -----------------------------------------------------------
    async def _on_startup(self, app) -> None:
        if self._auto_register:
            asyncio.ensure_future(self.register_with_manager())

    async def _on_cleanup(self, app) -> None:
        if self._heartbeat_manager is not None:
            await self._heartbeat_manager.stop()
        if self._round_task is not None and not self._round_task.done():
            self._round_task.cancel()
        if self._session is not None and not self._session.closed:
            await self._session.close()
        self._executor.shutdown(wait=False, cancel_futures=True)

    def rebind(self, app: web.Application) -> None:
        """Attach this worker (model, data plane and all) to a fresh aiohttp application after its previous HTTP
        front-end was shut down -- a seat coming back after a crash of its web process keeps its GPU state.  The caller
        starts the site and calls :meth:`register_with_manager`; the manager hands out a new client id."""
        self.app = app
        self.client_id = self.key = None
        self.update_in_progress = False
        self._round_task = None
        self._heartbeat_manager = None
        self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="baton-train",
                                            initializer=self._bind_device)
        self._auto_register = False
        self.register_handlers()
        app.on_cleanup.append(self._on_cleanup)

    def _get_session(self) -> aiohttp.ClientSession:
        if self._session is None or self._session.closed:
            self._session = aiohttp.ClientSession()
        return self._session

    def _auth_query(self) -> str:
        return "?client_id={}&key={}".format(self.client_id, self.key)

    # -- registration / heartbeat ---------------------------------------------
    @ensure_no_collision
    async def register_with_manager(self) -> bool:
        url = urljoin(self.manager_url, "register")
        data = {"url": self.worker_host, "port": self.port}
        data.update(self.plane.registration_extras())
        timeout = 1.0
        while True:
            try:
                async with self._get_session().get(url, json=data) as resp:
                    if resp.status == 200:
                        response = await resp.json()
                        self.client_id = response["client_id"]
                        self.key = response["key"]
                        break
                    log.warning("register got HTTP %d", resp.status)
            except aiohttp.ClientError:
                pass
            log.info("manager unreachable; retrying registration in %.0fs", timeout)
            await asyncio.sleep(timeout)
            timeout = min(timeout * 2, MAX_BACKOFF)
        log.info("registered as %s", self.client_id)
        if self._heartbeat_manager is not None:
            await self._heartbeat_manager.stop()
        self._heartbeat_manager = PeriodicTask(self.heartbeat, self.heartbeat_time).start()
        return True

    @ensure_no_collision
    async def heartbeat(self) -> bool:
        timeout = 1.0
        while True:
            url = urljoin(self.manager_url, "heartbeat")
            data = {"client_id": self.client_id, "key": self.key}
            try:
                async with self._get_session().get(url, json=data) as resp:
                    if resp.status == 200:
                        return True
                    if resp.status == 401:
                        log.info("manager forgot us; re-registering")
                        asyncio.ensure_future(self.register_with_manager())
                        return False
            except aiohttp.ClientError:
                pass
            log.info("could not reach manager; waiting %.0fs", timeout)
            await asyncio.sleep(timeout)
            timeout = min(timeout * 2, MAX_BACKOFF)

    def _bind_device(self) -> None:
        dev = getattr(self, "device", None)
        if dev is not None and getattr(dev, "type", None) == "cuda":
            import torch
            torch.cuda.set_device(dev)

    # -- routes ---------------------------------------------------------------
    def register_handlers(self) -> None:
        r = self.app.router
        r.add_post("/{}/round_start".format(self.name), self.round_start)
        r.add_post("/{}/aggregate".format(self.name), self.aggregate)
        r.add_get("/{}/state_dict".format(self.name), self.get_state_dict)

    def _credentials_ok(self, request: web.Request) -> bool:
        cid = request.query.get("client_id") or request.headers.get("X-Baton-Client-Id")
        key = request.query.get("key") or request.headers.get("X-Baton-Key")
        return self.client_id is not None and cid == self.client_id and key == self.key

    async def round_start(self, request: web.Request) -> web.Response:
        if self.update_in_progress:
            return web.json_response({"err": "Update in Progress"}, status=409)
        body = await request.read()
        if not self._credentials_ok(request):
            asyncio.ensure_future(self.register_with_manager())
            return web.json_response({"err": "Wrong Client"}, status=404)
        try:
            data = wire.loads(body, trusted=self.trusted_peers)
        except Exception as exc:
            log.warning("undecodable round_start: %r", exc)
            return web.json_response({"err": "Bad Payload"}, status=400)
        self.last_update = update_name = data["update_name"]
        n_epoch = int(data["n_epoch"])
        self.update_in_progress = True
        try:
            self.plane.receive_round(self, data)
        except Exception:
            self.update_in_progress = False
            log.exception("could not load round weights")
            return web.json_response({"err": "Bad Weights"}, status=400)
        self._round_task = asyncio.ensure_future(self._run_round(update_name, n_epoch))
        return web.json_response("OK")

    async def aggregate(self, request: web.Request) -> web.Response:
        """GPU-seated planes only: run this rank's share of the fused
        reduce+broadcast with the weight vector chosen by the manager."""
        if not self._credentials_ok(request):
            return web.json_response({"err": "Wrong Client"}, status=404)
        plan = wire.loads(await request.read(), trusted=self.trusted_peers)
        try:
            await run_blocking(self.plane.aggregate, self, plan, executor=self._executor)
        except NotImplementedError:
            return web.json_response({"err": "No Data Plane"}, status=501)
        return web.json_response("OK")

    async def get_state_dict(self, request: web.Request) -> web.Response:
        if not self._credentials_ok(request):
            return web.json_response({"err": "Wrong Client"}, status=404)
        body = await run_blocking(self.plane.export_state, self, executor=self._executor)
        return web.Response(body=body, content_type="application/octet-stream")

    # -- the round ---------------------------------------------------------------
    def _train_blocking(self, n_epoch: int):
        if self.fail_next_rounds > 0:
            self.fail_next_rounds -= 1
            raise RuntimeError("injected training failure")
        import time
        t0 = time.perf_counter()
        data, n_samples = self.get_data()
        train = resolve_train_fn(self.model)
        loss_history = train(*data, n_epoch=n_epoch, **self.train_kwargs)
        out = n_samples, [float(x) for x in loss_history]     # reading the losses waits for the device
        if getattr(self, "train_seconds", None) is not None:
            self.train_seconds.append(time.perf_counter() - t0)
        return out

    async def _run_round(self, update_name: str, n_epoch: int) -> None:
        try:
            n_samples, loss_history = await run_blocking(self._train_blocking, n_epoch,
                                                         executor=self._executor)
            self.last_loss_history = loss_history
            if self.drop_next_reports > 0:
                self.drop_next_reports -= 1
                log.info("injected fault: not reporting %s", update_name)
                return
            await self.report_update(update_name, n_samples, loss_history)
        except asyncio.CancelledError:
            raise
        except Exception:
            log.exception("round %s failed on %s", update_name, self.client_id)
        finally:
            self.update_in_progress = False

    async def report_update(self, update_name: str, n_samples: int, loss_history) -> int:
        url = urljoin(self.manager_url, "update") + self._auth_query()
        body = await run_blocking(self.plane.update_message, self, update_name, n_samples,
                                  loss_history, executor=self._executor)
        async with self._get_session().post(url, data=body) as resp:
            if resp.status == 200:
                self.n_updates += 1
            elif resp.status == 401:
                asyncio.ensure_future(self.register_with_manager())
            elif resp.status == 410:
                log.info("sent a stale update (%s)", update_name)
            return resp.status

    def get_data(self):
        """Return ``(args_tuple_for_train, n_samples)`` for this client's
        private shard.  Subclasses must implement it."""
        raise NotImplementedError
