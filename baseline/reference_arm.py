"""Reference arm of the benchmark: the UNMODIFIED reference (baseline/_ref, byte-identical copy
of mynameisfiber/baton) driven through its own public API and stock code path.

Nothing from ``baton_b200`` is imported here.  What runs:
  * ``manager.Manager`` / ``Experiment`` (reference manager.py) in rank 0, HTTP on localhost;
  * one ``worker.ExperimentWorker`` subclass per rank (reference worker.py) on that rank's GPU;
  * the user model = stock ``torchvision.models.resnet18(num_classes=10)`` with the reference's
    model contract (``name``, ``train(X, y, n_epoch)`` written like reference demo.py:29-49 with the
    reference's own ``utils.EpochProgress``), bf16 autocast, plain ``torch.optim.SGD``;
  * rounds are triggered with ``GET /{name}/start_round?n_epoch=E`` exactly as an operator would.

Per round the reference pickles the full state_dict to every worker over HTTP, each worker trains
and POSTs its full pickled state_dict back, and the manager reduces on the CPU (manager.py:113-126).
The only compatibility shim is ``collections.Iterator`` (removed in Python 3.10; reference
utils.py:5 imports it) -- an attribute set on the stdlib module, no reference file is touched.

Metric, config and JSON line are the same as the product arm (bench.py).
"""
from __future__ import annotations

import argparse
import asyncio
import collections
import collections.abc
import json
import os
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


_REAL_FD = None


def _unavailable(why: str) -> None:
    if int(os.environ.get("RANK", "0")) == 0:
        os.write(_REAL_FD if _REAL_FD is not None else 1, (json.dumps({"impl": "reference", "unavailable": why}) + "\n").encode())
    sys.exit(0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--local-epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--port-base", type=int, default=int(os.environ.get("BATON_REF_PORT", "18700")))
    ap.add_argument("--role", default="bench", choices=["bench", "manager"])
    args, _ = ap.parse_known_args()
    # the reference logs with print(); keep stdout clean for the single JSON result line
    sys.stdout.flush()
    global _REAL_FD
    _REAL_FD = os.dup(1)
    real_stdout = os.fdopen(os.dup(_REAL_FD), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    if not os.path.exists(os.path.join(REF, "manager.py")):
        _unavailable("baseline/_ref is empty: run baseline/install_reference.sh (reference is not pip-installable)")
    # the reference imports its siblings by bare module name: make baseline/_ref the ONLY candidate
    repo_root = os.path.dirname(HERE)
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (repo_root, HERE)]
    sys.path.insert(0, REF)
    collections.Iterator = collections.abc.Iterator  # py>=3.10 shim for reference utils.py:5

    try:
        import torch
        import torch.distributed as dist
        import torchvision
        from aiohttp import web
        import aiohttp
        import manager as ref_manager            # noqa: E402  (reference modules)
        import worker as ref_worker              # noqa: E402
        import utils as ref_utils                # noqa: E402
    except Exception as exc:  # pragma: no cover
        _unavailable("reference import failed: {!r}".format(exc))
    assert os.path.abspath(ref_manager.__file__).startswith(REF), ref_manager.__file__

    if args.role == "manager":
        # the parameter server: its own OS process, CPU only, exactly what reference demo.py:68-73,77 does
        import torchvision as tv

        class CpuModel(torch.nn.Module):
            name = "resnet18"

            def __init__(self):
                super().__init__()
                self.net = tv.models.resnet18(num_classes=10)

            def state_dict(self, *a, **kw):   # live tensors (the reference writes into them in place)
                return collections.OrderedDict((k, v) for k, v in self.net.state_dict().items() if v.dim() > 0)

        app = web.Application(client_max_size=1 << 32)
        ref_manager.Manager(app).register_experiment(CpuModel())
        web.run_app(app, host="127.0.0.1", port=args.port_base, print=None)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    on_gpu = torch.cuda.is_available()
    if not on_gpu and not os.environ.get("BATON_REF_ALLOW_CPU"):
        _unavailable("no CUDA device")
    dev = torch.device("cuda", local_rank) if on_gpu else torch.device("cpu")
    if on_gpu:
        torch.cuda.set_device(dev)
    if world > 1:
        # host-side barriers only (gloo): an NCCL barrier kernel parked on the GPU would block the
        # worker's training kernels, which the reference launches from its aiohttp handler thread
        dist.init_process_group("gloo")

    # ------------------------------------------------------------------ user model (reference contract)
    class Model(torch.nn.Module):
        name = "resnet18"

        def __init__(self, device):
            super().__init__()
            self.net = torchvision.models.resnet18(num_classes=10).to(device)
            self.device = device

        def forward(self, X):
            return self.net(X)

        def state_dict(self, *a, **kw):
            # tensors cross the wire as CPU tensors, as in the reference.  0-dim integer buffers
            # (BatchNorm num_batches_tracked) are left out: reference manager.py:126 does
            # ``value[:] = ...`` which raises IndexError on a 0-dim tensor and aborts the FedAvg loop.
            return collections.OrderedDict((k, v.detach().cpu()) for k, v in self.net.state_dict().items()
                                           if v.dim() > 0)

        def load_state_dict(self, sd, *a, **kw):
            return self.net.load_state_dict(sd, strict=False)

        def train(self, X=None, y=None, n_epoch=32, lr=args.lr, batch_size=args.batch_size, verbose=False):
            if X is None or isinstance(X, bool):
                return torch.nn.Module.train(self, True if X is None else X)
            torch.nn.Module.train(self, True)
            loss = torch.nn.CrossEntropyLoss()
            idxs = torch.randperm(X.shape[0], device=X.device)
            optimizer = torch.optim.SGD(self.net.parameters(), lr=lr)
            loss_history = []
            for epoch in range(n_epoch):
                batch_iter = ref_utils.EpochProgress(epoch, torch.split(idxs, batch_size), verbose=verbose)
                for batch_idxs in batch_iter:
                    optimizer.zero_grad()
                    X_batch = X[batch_idxs]
                    y_batch = y[batch_idxs]
                    with torch.autocast(dev.type, dtype=torch.bfloat16):
                        output = self(X_batch)
                    loss_batch = loss(output.float(), y_batch)
                    batch_iter.update_loss(loss_batch)     # float(loss): a host sync per batch, as in the reference
                    loss_batch.backward()
                    optimizer.step()
                loss_history.append(batch_iter.loss)
            return loss_history

    # synthetic non-IID shard of the named shape in pinned host memory (same generator family as the
    # product arm: class-conditional Gaussians, Dirichlet label skew), NCHW for the stock model
    g = torch.Generator().manual_seed(1234 + rank)
    probs = torch._standard_gamma(torch.full((10,), 0.5), generator=g)
    probs = probs / probs.sum()
    y_host = torch.multinomial(probs, args.samples, replacement=True, generator=g)
    means = torch.randn(10, 3, 32, 32, generator=torch.Generator().manual_seed(7)) * 0.5
    X_host = (means[y_host] + torch.randn(args.samples, 3, 32, 32, generator=g)).to(torch.bfloat16)
    if on_gpu:
        X_host, y_host = X_host.pin_memory(), y_host.pin_memory()
    h2d_bytes = X_host.numel() * 2 + y_host.numel() * 8

    class Worker(ref_worker.ExperimentWorker):
        def get_data(self):
            X = X_host.to(dev, non_blocking=True).float()
            y = y_host.to(dev, non_blocking=True)
            return (X, y), args.samples

    mport = args.port_base
    wport = args.port_base + 1 + rank
    state = {}

    manager_proc = None

    def start_manager():
        import subprocess
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        return subprocess.Popen([sys.executable, os.path.abspath(__file__), "--role", "manager",
                                 "--port-base", str(mport)], env=env, stdout=subprocess.DEVNULL)

    async def manager_get(path):
        async with aiohttp.ClientSession() as s:
            async with s.get("http://127.0.0.1:{}/resnet18/{}".format(mport, path)) as r:
                return r.status, await r.json()

    async def start_worker():
        app = web.Application(client_max_size=1 << 32)
        w = Worker(app, Model(dev), "127.0.0.1:{}".format(mport), port=wport, heartbeat_time=600,
                   worker_host="http://127.0.0.1:{}/resnet18/".format(wport))
        runner = web.AppRunner(app)
        await runner.setup()
        await web.TCPSite(runner, "127.0.0.1", wport).start()
        state["worker"] = w

    loop = asyncio.new_event_loop()

    def run_loop():
        asyncio.set_event_loop(loop)
        loop.run_forever()

    threading.Thread(target=run_loop, daemon=True).start()

    def call(coro, timeout=600):
        return asyncio.run_coroutine_threadsafe(coro, loop).result(timeout)

    def barrier():
        if world > 1:
            dist.barrier()
        if on_gpu:
            torch.cuda.synchronize()

    if rank == 0:
        manager_proc = start_manager()
        t0 = time.time()
        while time.time() - t0 < 120:
            try:
                if call(manager_get("clients"))[0] == 200:
                    break
            except Exception:
                time.sleep(0.2)
    barrier()
    call(start_worker())
    # wait until every worker has registered with the manager
    t0 = time.time()
    while state["worker"].client_id is None and time.time() - t0 < 60:
        time.sleep(0.05)
    barrier()
    done = {"rounds": 0}

    async def one_round(n_epoch):
        status, accepted = await manager_get("start_round?n_epoch={}".format(n_epoch))
        assert status == 200 and len(accepted) == world and all(accepted.values()), (status, accepted)
        done["rounds"] += 1
        while True:   # /clients is the reference's working introspection endpoint (client_manager.py:139-142)
            _, clients = await manager_get("clients")
            if all(c["num_updates"] >= done["rounds"] for c in clients):
                return None
            await asyncio.sleep(0.002)

    def rounds(k):
        last = None
        for _ in range(k):
            if rank == 0:
                last = call(one_round(args.local_epochs), timeout=900)
            barrier()
        return last

    try:
        rounds(args.warmup)
        barrier()
        t0 = time.perf_counter()
        last_loss = rounds(args.steps)
        barrier()
        dt = time.perf_counter() - t0
    except BaseException as exc:   # a wedged reference round must not hang the driver
        if manager_proc is not None:
            manager_proc.terminate()
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "reference round failed: {!r}".format(exc)}),
                  file=real_stdout, flush=True)
        os._exit(0)
    t = torch.tensor([dt])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        total_samples = world * args.samples * args.local_epochs * args.steps
        value = total_samples / dt
        n_param_bytes = sum(v.numel() * v.element_size() for v in state["worker"].model.state_dict().values())
        print(json.dumps({
            "impl": "reference",
            "metric": "federated local samples/sec (whole box), ResNet-18 FedAvg, synthetic non-IID 32x32 shards",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "rounds_per_s": args.steps / dt, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (autocast)", "data": "synthetic",
            "timing": "wall clock on rank 0 bracketed by barrier+synchronize, max over ranks "
                      "(the reference's data plane is host-side HTTP+pickle: there is no device timeline to time)",
            "config": {"model": "resnet18(num_classes=10)", "global_batch": world * args.batch_size,
                       "batch_size": args.batch_size, "samples_per_client": args.samples, "image": "32x32x3",
                       "local_epochs": args.local_epochs, "parallelism": "fedavg dp{}".format(world),
                       "transport": "HTTP/1.1 + pickle via aiohttp (reference stock path)",
                       "state_dict_bytes": n_param_bytes},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": n_param_bytes},
            "final_loss": last_loss,
        }), file=real_stdout, flush=True)
    if manager_proc is not None:
        manager_proc.terminate()
    if world > 1:
        dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
