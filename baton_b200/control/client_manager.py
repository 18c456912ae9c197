"""Client registry, authentication, liveness and fan-out notification.

Parity target: ``ClientManager`` (reference client_manager.py:14-150).

Wire protocol kept byte-compatible:
  GET /{name}/register   JSON body {"url": str|null, "port": int}
                         -> 200 {"client_id", "key"}            (:86-111)
  GET /{name}/heartbeat  JSON body {"client_id","key"} -> 200 "OK" |
                         401 {"err": "Invalid Client"|"Invalid Key"} (:113-127)
  GET /{name}/clients    -> 200 JSON list of records, ``key`` stripped (:139-142)
Record schema (:100-109): key, client_id, remote, port, last_heartbeat, url,
last_update, num_updates.  Extra optional registration fields (``rank``,
``device``, ``backend``) describe the client's seat on the NVLink data plane.

Liveness (:129-137): a client whose last heartbeat is older than ``client_ttl``
is culled, periodically every ``client_ttl // 2`` seconds and before each
fan-out (:38).  A client is evicted on connect error or HTTP 404 during notify
(:58-61).

New: seeded client sampling (``sample``), eviction callbacks, a clock seam,
periodic task started from aiohttp ``on_startup`` and the HTTP session closed
on ``on_cleanup`` (quirks 15, 18).
"""
from __future__ import annotations

import asyncio
import logging
import random
from datetime import timedelta
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple
from urllib.parse import urljoin

import aiohttp
from aiohttp import web

from ..utils.aio import PeriodicTask
from ..utils.misc import SYSTEM_CLOCK, Clock, json_clean, random_key

log = logging.getLogger("baton_b200.clients")

EvictCallback = Callable[[str, str], None]


class ClientManager:
    def __init__(self, name: str, app: web.Application, client_ttl: float = 300,
                 *, clock: Clock = SYSTEM_CLOCK, seed: Optional[int] = None,
                 notify_timeout: Optional[float] = None):
        self.name = name
        self.app = app
        self.clock = clock
        self.client_ttl = timedelta(seconds=client_ttl)
        self.clients: Dict[str, dict] = {}
        self.notify_timeout = notify_timeout
        self._session: Optional[aiohttp.ClientSession] = None
        self._evict_callbacks: List[EvictCallback] = []
        self._rng = random.Random(seed)
        self.n_registered = 0
        self.n_evicted = 0
        self.register_handlers()
        self._stale_manager = PeriodicTask(self.cull_clients, max(client_ttl / 2.0, 0.05))
        app.on_startup.append(self._on_startup)
        app.on_cleanup.append(self._on_cleanup)

    # -- aiohttp lifecycle -------------------------------------------------
    async def _on_startup(self, app) -> None:
        self._stale_manager.start()

    async def _on_cleanup(self, app) -> None:
        await self._stale_manager.stop()
        if self._session is not None and not self._session.closed:
            await self._session.close()
        self._session = None

    def _get_session(self) -> aiohttp.ClientSession:
        if self._session is None or self._session.closed:
            timeout = aiohttp.ClientTimeout(total=self.notify_timeout)
            self._session = aiohttp.ClientSession(timeout=timeout)
        return self._session

    # -- dict-like access (client_manager.py:26-27,80-84) ------------------
    def __len__(self) -> int:
        return len(self.clients)

    def __contains__(self, client_id: str) -> bool:
        return client_id in self.clients

    def __getitem__(self, key: str) -> dict:
        return self.clients[key]

    def __setitem__(self, key: str, value: dict) -> None:
        self.clients[key] = value

    def __iter__(self):
        return iter(self.clients)

    # -- eviction ----------------------------------------------------------
    def add_evict_callback(self, cb: EvictCallback) -> None:
        self._evict_callbacks.append(cb)

    def evict(self, client_id: str, reason: str) -> Optional[dict]:
        rec = self.clients.pop(client_id, None)
        if rec is not None:
            self.n_evicted += 1
            log.info("evicting %s (%s)", client_id, reason)
            for cb in self._evict_callbacks:
                try:
                    cb(client_id, reason)
                except Exception:  # pragma: no cover
                    log.exception("evict callback failed")
        return rec

    # -- sampling (new; the reference notifies everyone, :39-44) -----------
    def sample(self, k: Optional[int] = None, fraction: Optional[float] = None,
               among: Optional[Iterable[str]] = None) -> List[str]:
        """Pick the participants of a round.  ``k=None`` and ``fraction=None``
        reproduces the reference (all live clients).  The draw is seeded so a
        run is reproducible."""
        pool = sorted(self.clients if among is None else [c for c in among if c in self.clients])
        if fraction is not None:
            k = max(1, int(round(fraction * len(pool)))) if pool else 0
        if k is None or k >= len(pool):
            return pool
        if k <= 0:
            return []
        return sorted(self._rng.sample(pool, k))

    # -- fan-out -----------------------------------------------------------
    async def notify_clients(self, client_method: str, http_method: str = "GET",
                             client_callback=None, notify_callback=None,
                             clients: Optional[Sequence[str]] = None,
                             per_client_kwargs: Optional[Callable[[str], dict]] = None,
                             **kwargs) -> List[Tuple[str, bool]]:
        """Concurrently call ``{client.url}{client_method}`` on every client (or
        on ``clients``) and return ``[(client_id, ok)]`` (:35-47)."""
        await self.cull_clients()
        targets = list(self.clients) if clients is None else [c for c in clients if c in self.clients]
        coros = []
        for c in targets:
            kw = dict(kwargs)
            if per_client_kwargs is not None:
                kw.update(per_client_kwargs(c))
            coros.append(self.notify_client(c, client_method, http_method=http_method,
                                            callback=client_callback, **kw))
        result = list(await asyncio.gather(*coros)) if coros else []
        if notify_callback is not None:
            return await notify_callback(result)
        return result

    async def notify_client(self, client_id: str, client_method: str,
                            http_method: str = "GET", callback=None, **kwargs) -> Tuple[str, bool]:
        rec = self.clients.get(client_id)
        if rec is None:
            return client_id, False
        url = urljoin(rec["url"], client_method)
        # credentials ride in the query string for wire compatibility (:52) and
        # are mirrored in headers for peers that prefer them.
        url += "?client_id={}&key={}".format(client_id, rec["key"])
        headers = dict(kwargs.pop("headers", None) or {})
        headers.setdefault("X-Baton-Client-Id", client_id)
        headers.setdefault("X-Baton-Key", rec["key"])
        result = False
        try:
            async with self._get_session().request(http_method, url, headers=headers, **kwargs) as resp:
                if resp.status == 200:
                    result = True
                elif resp.status == 404:
                    self.evict(client_id, "404 on notify")
                await resp.read()
        except aiohttp.ClientConnectorError:
            self.evict(client_id, "connect error")
        except (aiohttp.ClientError, asyncio.TimeoutError) as exc:
            log.warning("notify %s failed: %r", client_id, exc)
        if callback is not None:
            await callback(client_id, result)
        return client_id, result

    # -- routes ------------------------------------------------------------
    def register_handlers(self) -> None:
        r = self.app.router
        r.add_get("/{}/register".format(self.name), self.register)
        r.add_get("/{}/clients".format(self.name), self.get_clients)
        r.add_get("/{}/heartbeat".format(self.name), self.heartbeat)

    async def register(self, request: web.Request) -> web.Response:
        try:
            data = await request.json()
        except Exception:
            return web.json_response({"err": "Invalid Body"}, status=400)
        if not isinstance(data, dict) or ("port" not in data and not data.get("url")):
            return web.json_response({"err": "Missing port"}, status=400)
        remote = request.remote
        client_id = "client_{}_{}".format(self.name, random_key(6))
        while client_id in self.clients:  # 52**6 ids; collisions are possible in principle
            client_id = "client_{}_{}".format(self.name, random_key(6))
        key = random_key()
        if data.get("url"):
            url = data["url"]
            if not url.endswith("/"):
                url += "/"
        else:
            url = "http://{}:{}/{}/".format(remote, data["port"], self.name)
        rec = {
            "key": key,
            "client_id": client_id,
            "remote": remote,
            "port": data.get("port"),
            "last_heartbeat": self.clock.now(),
            "url": url,
            "last_update": None,
            "num_updates": 0,
        }
        for extra in ("rank", "device", "backend", "logical_clients"):
            if extra in data:
                rec[extra] = data[extra]
        self.clients[client_id] = rec
        self.n_registered += 1
        log.info("registered %s at %s", client_id, url)
        return web.json_response({"client_id": client_id, "key": key})

    async def heartbeat(self, request: web.Request) -> web.Response:
        try:
            data = await request.json()
            client_id, key = data["client_id"], data["key"]
        except Exception:
            return web.json_response({"err": "Invalid Client"}, status=401)
        rec = self.clients.get(client_id)
        if rec is None:
            return web.json_response({"err": "Invalid Client"}, status=401)
        if rec["key"] != key:
            return web.json_response({"err": "Invalid Key"}, status=401)
        rec["last_heartbeat"] = self.clock.now()
        return web.json_response("OK")

    async def cull_clients(self) -> List[str]:
        now = self.clock.now()
        stale = [cid for cid, rec in self.clients.items()
                 if (now - rec["last_heartbeat"]) > self.client_ttl]
        for cid in stale:
            self.evict(cid, "stale heartbeat")
        return stale

    async def get_clients(self, request: web.Request) -> web.Response:
        return web.json_response([json_clean(rec) for rec in self.clients.values()])

    def verify_request(self, request: web.Request) -> str:
        """Authenticate a worker call (:144-150).  Query-string credentials
        first, ``X-Baton-*`` headers as a fallback; 401 on any mismatch."""
        client_id = request.query.get("client_id") or request.headers.get("X-Baton-Client-Id")
        client_key = request.query.get("key") or request.headers.get("X-Baton-Key")
        rec = self.clients.get(client_id) if client_id else None
        if rec is None or client_key != rec["key"]:
            raise web.HTTPUnauthorized()
        return client_id
