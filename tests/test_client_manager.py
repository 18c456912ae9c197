import asyncio

from aiohttp import web
from aiohttp.test_utils import TestClient, TestServer

from baton_b200.control import ClientManager
from baton_b200.utils import FakeClock
from conftest import run_async


async def _mk(ttl=300, clock=None, seed=None):
    app = web.Application()
    cm = ClientManager("exp", app, ttl, clock=clock or FakeClock(), seed=seed)
    server = TestServer(app, host="127.0.0.1")
    await server.start_server()
    return cm, TestClient(server), server


async def _register(client, **body):
    body.setdefault("port", 1234)
    body.setdefault("url", None)
    async with client.get("/exp/register", json=body) as r:
        assert r.status == 200
        return await r.json()


@run_async
async def test_register_heartbeat_clients_schema():
    cm, client, server = await _mk()
    try:
        cred = await _register(client, port=4321)
        assert set(cred) == {"client_id", "key"}
        assert cred["client_id"].startswith("client_exp_") and len(cred["client_id"]) == len("client_exp_") + 6
        assert len(cred["key"]) == 32
        rec = cm[cred["client_id"]]
        assert set(rec) >= {"key", "client_id", "remote", "port", "last_heartbeat", "url",
                            "last_update", "num_updates"}
        assert rec["url"] == "http://127.0.0.1:4321/exp/" and rec["num_updates"] == 0
        cred2 = await _register(client, url="http://example.test:9/exp")
        assert cm[cred2["client_id"]]["url"] == "http://example.test:9/exp/"
        async with client.get("/exp/heartbeat", json=cred) as r:
            assert r.status == 200 and await r.json() == "OK"
        async with client.get("/exp/heartbeat", json={"client_id": "nope", "key": "k"}) as r:
            assert r.status == 401 and (await r.json())["err"] == "Invalid Client"
        async with client.get("/exp/heartbeat", json={"client_id": cred["client_id"], "key": "bad"}) as r:
            assert r.status == 401 and (await r.json())["err"] == "Invalid Key"
        async with client.get("/exp/clients") as r:
            data = await r.json()
        assert len(data) == 2 and all("key" not in d for d in data)
        assert isinstance(data[0]["last_heartbeat"], str)
        async with client.get("/exp/register", data=b"not json") as r:
            assert r.status == 400
    finally:
        await client.close(); await server.close()


@run_async
async def test_ttl_culling_with_fake_clock():
    clock = FakeClock()
    cm, client, server = await _mk(ttl=300, clock=clock)
    try:
        a = await _register(client)
        clock.advance(200)
        b = await _register(client)
        clock.advance(150)                       # a: 350 s stale, b: 150 s
        evicted = []
        cm.add_evict_callback(lambda cid, why: evicted.append((cid, why)))
        stale = await cm.cull_clients()
        assert stale == [a["client_id"]] and b["client_id"] in cm and len(cm) == 1
        assert evicted == [(a["client_id"], "stale heartbeat")]
        async with client.get("/exp/heartbeat", json=b) as r:
            assert r.status == 200
        clock.advance(299)
        assert await cm.cull_clients() == []     # heartbeat refreshed the TTL
    finally:
        await client.close(); await server.close()


@run_async
async def test_verify_request_and_headers():
    cm, client, server = await _mk()
    try:
        cred = await _register(client)

        async def probe(request):
            return web.json_response(cm.verify_request(request))
        # routes are frozen after start; use a fresh app to exercise verify_request
        app2 = web.Application()
        app2.router.add_get("/p", probe)
        s2 = TestServer(app2, host="127.0.0.1"); await s2.start_server(); c2 = TestClient(s2)
        async with c2.get("/p", params={"client_id": cred["client_id"], "key": cred["key"]}) as r:
            assert r.status == 200 and await r.json() == cred["client_id"]
        async with c2.get("/p", params={"client_id": cred["client_id"], "key": "bad"}) as r:
            assert r.status == 401
        async with c2.get("/p") as r:
            assert r.status == 401
        async with c2.get("/p", headers={"X-Baton-Client-Id": cred["client_id"], "X-Baton-Key": cred["key"]}) as r:
            assert r.status == 200
        await c2.close(); await s2.close()
    finally:
        await client.close(); await server.close()


@run_async
async def test_sampling_is_seeded_and_bounded():
    cm, client, server = await _mk(seed=7)
    cm2, client2, server2 = await _mk(seed=7)
    try:
        for _ in range(16):
            await _register(client)
        assert len(cm.sample(None)) == 16
        assert len(cm.sample(100)) == 16
        assert cm.sample(0) == []
        picks = cm.sample(4)
        assert len(picks) == 4 and len(set(picks)) == 4 and set(picks) <= set(cm.clients)
        assert len(cm.sample(fraction=0.25)) == 4
        # same seed + same pool -> same draw
        cm2.clients = dict(cm.clients)
        cmA = ClientManager.__new__(ClientManager)
        import random
        cm._rng = random.Random(3); cm2._rng = random.Random(3)
        assert cm.sample(5) == cm2.sample(5)
    finally:
        await client.close(); await server.close(); await client2.close(); await server2.close()


@run_async
async def test_notify_evicts_on_404_and_connect_error():
    cm, client, server = await _mk()
    hits = []

    async def ok(request):
        hits.append(dict(request.query))
        return web.json_response("OK")

    async def gone(request):
        return web.json_response({"err": "Wrong Client"}, status=404)

    async def busy(request):
        return web.json_response({"err": "busy"}, status=409)

    wapp = web.Application()
    wapp.router.add_post("/ok/ping", ok)
    wapp.router.add_post("/gone/ping", gone)
    wapp.router.add_post("/busy/ping", busy)
    ws = TestServer(wapp, host="127.0.0.1"); await ws.start_server()
    try:
        base = "http://127.0.0.1:{}".format(ws.port)
        a = await _register(client, url=base + "/ok/")
        b = await _register(client, url=base + "/gone/")
        c = await _register(client, url=base + "/busy/")
        d = await _register(client, url="http://127.0.0.1:1/dead/")
        seen = []

        async def cb(cid, res):
            seen.append((cid, res))
        result = dict(await cm.notify_clients("ping", http_method="POST", data=b"x", client_callback=cb))
        assert result == {a["client_id"]: True, b["client_id"]: False, c["client_id"]: False, d["client_id"]: False}
        assert set(cm.clients) == {a["client_id"], c["client_id"]}   # 404 + connect error evicted, 409 kept
        assert hits[0]["client_id"] == a["client_id"] and hits[0]["key"] == a["key"]
        assert len(seen) == 4
        only = await cm.notify_clients("ping", http_method="POST", clients=[a["client_id"], "ghost"])
        assert only == [(a["client_id"], True)]
    finally:
        await client.close(); await server.close(); await ws.close()
