"""CPU / gloo worker for tests/test_engine_gloo.py (torchrun --nproc-per-node 2 tests/mp_engine_gloo.py).

Drives the SPMD :class:`FederatedEngine` -- the same class the GPU bench uses -- through the
``torch.distributed`` session on gloo: plain rounds (every rank one client) and the client-sampling
configuration (logical clients time-sliced over the ranks, k sampled per round), and checks the global
model against the closed-form FedAvg of the per-client results."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baton_b200.models import MLP2  # noqa: E402
from baton_b200.parallel.engine import FederatedEngine  # noqa: E402


def shard(cid, n):
    g = torch.Generator().manual_seed(1000 + cid)
    X = torch.randn(n, 10, generator=g)
    w = torch.arange(1, 11, dtype=torch.float32)
    return X, (X @ w).unsqueeze(1) + 0.01 * torch.randn(n, 1, generator=g)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fails = []

    def expect(cond, msg):
        ok = torch.tensor([1 if cond else 0])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            fails.append(msg)
        if rank == 0:
            print(("ok   " if int(ok) else "FAIL ") + msg, flush=True)

    # ---- plain rounds: rank r is client r with n_r samples -------------------------------------------
    torch.manual_seed(0)
    eng = FederatedEngine(MLP2(10, 16, 1), "cpu", backend="nccl", loss="mse", lr=0.01, batch_size=16,
                          wire_dtype="fp32", name="gloo")
    sizes = [32 * (r + 1) for r in range(world)]
    X, y = shard(rank, sizes[rank])
    g0 = eng.arena.global_w.clone()
    res = eng.run_round((X, y), n_epoch=2)
    # closed form: every rank's local result, gathered, weighted by n_k
    # (the engine already averaged, so recompute the local models from an identical second engine)
    torch.manual_seed(0)
    twin = FederatedEngine(MLP2(10, 16, 1), "cpu", backend="nccl", loss="mse", lr=0.01, batch_size=16,
                           wire_dtype="fp32", name="twin")
    expect(torch.equal(twin.arena.global_w, g0), "replicas start identical")
    expect(res.n_samples == sizes[rank] and len(res.loss_history) == 2, "round result carries n_k and per-epoch losses")
    same = [torch.empty_like(eng.arena.theta) for _ in range(world)]
    dist.all_gather(same, eng.arena.theta.clone())
    expect(all(torch.equal(same[0], t) for t in same), "global model identical on every rank after the round")
    expect(not torch.equal(eng.arena.theta, g0), "the round moved the global model")
    expect(torch.equal(eng.arena.theta, eng.arena.global_w), "theta == frozen global copy after the broadcast")
    gl = eng.global_loss(2)
    expect(len(gl) == 2 and gl[1] < gl[0], "sample-weighted loss history decreases across local epochs")
    first = gl[0]
    for _ in range(6):
        eng.run_round((X, y), n_epoch=2)
    expect(eng.global_loss(2)[1] < 0.5 * first, "loss keeps falling over rounds")

    # ---- client sampling: 6 logical clients over `world` ranks, 3 sampled per round -------------------
    torch.manual_seed(1)
    eng2 = FederatedEngine(MLP2(10, 16, 1), "cpu", backend="nccl", loss="mse", lr=0.01, batch_size=16,
                           wire_dtype="fp32", logical_clients=6, sample_k=3, seed=5, name="sampled")
    shards = {c: shard(c, 16 * (1 + c % 3)) for c in range(6)}
    seen = set()
    for r in range(5):
        before = eng2.arena.global_w.clone()
        res = eng2.run_round(lambda cid: shards[cid], n_epoch=1)
        parts = res.participants
        seen.update(parts)
        expect(len(parts) == 3 and parts == sorted(parts), "round {}: 3 sampled logical clients".format(r))
        plist = [None] * world
        dist.all_gather_object(plist, parts)
        expect(all(p == plist[0] for p in plist), "round {}: every rank drew the same participants".format(r))
        mine = [c for c in parts if c % world == rank]
        expect(res.n_samples == sum(shards[c][0].shape[0] for c in mine), "round {}: n_k = hosted samples".format(r))
        tot = torch.tensor([float(res.n_samples)])
        dist.all_reduce(tot)
        expect(int(tot) == sum(shards[c][0].shape[0] for c in parts), "round {}: weights cover all participants".format(r))
        expect(not torch.equal(before, eng2.arena.global_w), "round {}: global model updated".format(r))
    expect(len(seen) > 3, "sampling visits different clients across rounds")

    dist.barrier()
    if rank == 0:
        print("RESULT", "FAIL" if fails else "PASS", len(fails), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
