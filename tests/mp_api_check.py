"""Multi-GPU check of Baton's API on the NVLink data plane (launched by torchrun from test_gpu_fedavg.py or by hand:
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp_api_check.py).

One CPU ``Manager`` process + one ``GpuExperimentWorker`` per GPU over HTTP, rounds triggered by ``GET /start_round``:
  1. seats start from DIFFERENT random weights; after round 1 every replica equals the manager's model + the same
     update (the manager distributes its model with the first round_start) -> replicas bit-identical;
  2. a seat is killed (its HTTP site goes away): the next round evicts it, the survivors aggregate without it -- no hang;
  3. the seat comes back (new client id): it is handed the global model again, re-enters with the manager's barrier
     epoch, and the replicas are bit-identical again.
"""
import asyncio
import os
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import aiohttp  # noqa: E402
from aiohttp import web  # noqa: E402

from baton_b200.control.gpu_worker import GpuExperimentWorker  # noqa: E402
from baton_b200.data import dirichlet_label_shards, image_shard  # noqa: E402
from baton_b200.models import resnet18  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    host = dist.new_group(backend="gloo")
    name = "resnet18"
    mport = 18000 + int(os.environ.get("MASTER_PORT", "29500")) % 1000
    wport = mport + 1 + rank
    fails = []

    def ok(cond, what):
        if rank == 0:
            print(("ok   " if cond else "FAIL ") + what, flush=True)
        if not cond:
            fails.append(what)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier(group=host)

    torch.manual_seed(100 + rank)                      # different weights per seat on purpose
    model = resnet18(10)
    specs = dirichlet_label_shards(world, 10, 512, alpha=0.5, seed=11)
    X, y = image_shard(specs[rank], seed=3, dtype=torch.bfloat16, pin=True)
    mproc = None
    if rank == 0:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        mproc = subprocess.Popen([sys.executable, os.path.join(ROOT, "demo.py"), "manager", "x", str(mport), "--bind",
                                  "127.0.0.1", "--model", name, "--backend", "fused"], env=env, stdout=subprocess.DEVNULL)
    loop = asyncio.new_event_loop()
    threading.Thread(target=lambda: (asyncio.set_event_loop(loop), loop.run_forever()), daemon=True).start()

    def call(coro, timeout=300):
        return asyncio.run_coroutine_threadsafe(coro, loop).result(timeout)

    async def mget(path):
        async with aiohttp.ClientSession() as s:
            async with s.get("http://127.0.0.1:{}/{}/{}".format(mport, name, path)) as r:
                return r.status, await r.json()

    if rank == 0:
        t0 = time.time()
        while time.time() - t0 < 120:
            try:
                if call(mget("clients"))[0] == 200:
                    break
            except Exception:
                time.sleep(0.2)
    barrier()
    st = {}

    async def start_site():
        runner = web.AppRunner(st["app"])
        await runner.setup()
        site = web.TCPSite(runner, "127.0.0.1", wport)
        await site.start()
        st["runner"] = runner

    async def start_worker():
        st["app"] = web.Application(client_max_size=1 << 34)
        st["w"] = GpuExperimentWorker(st["app"], model, "127.0.0.1:{}".format(mport), device=dev,
                                      shard_fn=lambda: (X, y), backend="fused", port=wport, heartbeat_time=600,
                                      worker_host="http://127.0.0.1:{}/{}/".format(wport, name), n_ctas=64,
                                      train_kwargs={"lr": 0.05, "batch_size": 128})
        await start_site()

    call(start_worker(), timeout=600)
    w = st["w"]

    def cs(tag):
        a_ = w.arena
        print("   [rank {}] {}: theta {:.6f} global {:.6f} epoch {} rounds {} stale {}".format(
            rank, tag, float(a_.theta[: a_.n].double().sum()), float(a_.global_w[: a_.n].double().sum()),
            w.fed_session.epoch, w.fed_session.rounds, w.fed_session.stale), flush=True)
    _rr, _ag = w.plane.receive_round, w.plane.aggregate

    def rr(worker, msg):
        _rr(worker, msg)
        torch.cuda.synchronize()
        cs("after receive_round (state_dict in msg: {})".format("state_dict" in msg))

    def ag(worker, plan):
        torch.cuda.synchronize()
        cs("before aggregate plan round {} alive {}".format(plan.get("round"), plan.get("alive_ranks")))
        _ag(worker, plan)
        torch.cuda.synchronize()
        cs("after aggregate")
    if os.environ.get("API_CHECK_VERBOSE") == "1":
        w.plane.receive_round, w.plane.aggregate = rr, ag
    t0 = time.time()
    while w.client_id is None and time.time() - t0 < 60:
        time.sleep(0.05)
    barrier()
    rounds = {"n": 0}

    async def one_round(expect):
        status, accepted = await mget("start_round?n_epoch=1")
        assert status == 200, (status, accepted)
        rounds["n"] += 1
        t0 = time.time()
        while time.time() - t0 < 120:
            _, s_ = await mget("state")
            if not s_["in_progress"] and s_["n_updates"] >= rounds["n"]:
                return sum(1 for v in accepted.values() if v)
            await asyncio.sleep(0.005)
        raise TimeoutError("round did not close")

    def run_round(expect):
        n_ok = None
        if rank == 0:
            n_ok = call(one_round(expect), timeout=300)
        barrier()
        return n_ok

    def checksum():
        return w.arena.theta[: w.arena.n].double().sum().reshape(1).cpu()

    def diagnose(tag):
        """Which arena slots differ between rank 0 and the last rank (printed by rank 0)."""
        mine = torch.cat([w.arena.theta[: w.arena.n].cpu(), w.arena.global_w[: w.arena.n].cpu()])
        other = mine.clone()
        dist.broadcast(other, src=world - 1, group=host)
        if rank == 0:
            n = w.arena.n
            for nm, (a_, b_) in (("theta", (mine[:n], other[:n])), ("global_w", (mine[n:], other[n:]))):
                bad = [(name, float((a_[sl.offset: sl.offset + sl.numel] - b_[sl.offset: sl.offset + sl.numel]).abs().max()))
                       for name, sl in w.arena.slots.items()]
                bad = [x for x in bad if x[1] > 0]
                print("   diag[{}] {}: {} of {} slots differ, first {}; epoch {} rounds {}".format(
                    tag, nm, len(bad), len(w.arena.slots), bad[:3], w.fed_session.epoch, w.fed_session.rounds), flush=True)

    def identical(group_ranks):
        c = checksum()
        lo, hi = c.clone(), c.clone()
        if rank not in group_ranks:          # outsiders contribute neutral elements
            lo.fill_(float("inf")); hi.fill_(float("-inf"))
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=host)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=host)
        return float(lo) == float(hi)

    before = checksum()
    spread = [before.clone() for _ in range(world)]
    dist.all_gather(spread, before, group=host)
    ok(len({float(t) for t in spread}) == world, "seats start from different weights")
    n_ok = run_round(world)
    ok(rank != 0 or n_ok == world, "round 1: every seat accepted")
    ok(identical(range(world)), "round 1: replicas bit-identical after the manager's model was distributed")
    run_round(world)
    ok(identical(range(world)), "round 2: replicas bit-identical (metadata-only round)")

    victim = world - 1
    if rank == victim:                                  # the seat dies: its HTTP site disappears
        call(st["runner"].cleanup())
    barrier()
    n_ok = run_round(world - 1)
    ok(rank != 0 or n_ok == world - 1, "round 3: the dead seat was evicted, the survivors accepted")
    ok(identical(range(world - 1)), "round 3: survivors aggregated without the dead seat (no hang) and agree")
    if rank != victim:
        w.fed_session.check()
    if rank == victim:                                  # ... and comes back under a new client id
        async def revive():
            st["app"] = web.Application(client_max_size=1 << 34)
            w.rebind(st["app"])
            await start_site()
            await w.register_with_manager()
        call(revive(), timeout=120)
    barrier()
    time.sleep(0.5)
    n_ok = run_round(world)
    ok(rank != 0 or n_ok == world, "round 4: the returning seat was accepted again")
    same = identical(range(world))
    ok(same, "round 4: returning seat resynchronised (model + barrier epoch), replicas bit-identical")
    if not same:
        diagnose("round 4")
        run_round(world)
        ok(identical(range(world)), "round 5: replicas bit-identical")
        diagnose("round 5")
    flag = torch.tensor([len(fails)], dtype=torch.float64)
    dist.all_reduce(flag, group=host)
    if rank == 0:
        print("RESULT {} {}".format("PASS" if float(flag) == 0 else "FAIL", int(flag)), flush=True)
    if mproc is not None:
        mproc.terminate()
    os._exit(0 if float(flag) == 0 else 1)


if __name__ == "__main__":
    main()
