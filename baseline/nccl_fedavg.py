"""NCCL + cuBLAS/cuDNN baseline ("the reference's own NCCL build", BASELINE.json / BASELINE.md).

A faithful re-expression of the reference ALGORITHM with stock components only -- nothing from
``baton_b200`` is imported:
  * workers are one process per GPU, each a full replica of stock ``torchvision`` ResNet-18
    (cuDNN convolutions, cuBLAS GEMMs, bf16 autocast, channels_last, ``torch.optim.SGD``);
  * a round = ``local_epochs`` epochs of minibatch SGD over the private shard (loop shaped like
    reference demo.py:29-49; the running loss is accumulated on the device -- kinder than the
    reference's per-batch ``float(loss)``), then full-weight upload + sample-weighted mean over
    EVERY state_dict entry (manager.py:119-126) + full-state broadcast (manager.py:77-86), done as
    ONE NCCL all-reduce of the flattened, n_k/N-prescaled state in bf16 (same wire bytes as the
    fused kernel) followed by the local overwrite -- all clients participate (manager.py:82-89).

``--graph`` is the STRONG variant the judge asked for ("make the baseline honest"): the whole local
epoch (32 steps of index_select -> channels_last bf16 autocast forward -> loss -> backward -> SGD) is captured in ONE
CUDA graph, every float state_dict entry lives in one flat fp32 buffer (parameters and BatchNorm statistics are
views), and the round-end aggregate is scale -> bf16 cast -> ONE NCCL all-reduce -> copy back: three element-wise
kernels and one collective, no Python per-tensor loop.  The eager variant (~400 launches per step from Python, per-tensor
``torch.cat`` / copy aggregate) stays the default and is reported beside it.

Same metric / config / JSON contract as bench.py; ``"impl": "baseline"``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--local-epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--wire", default="bf16")
    ap.add_argument("--graph", action="store_true", help="CUDA-graph the local epoch + flat-buffer aggregate")
    args, _ = ap.parse_known_args()

    sys.stdout.flush()
    real_stdout = os.dup(1)      # keep stdout for the single JSON line; NCCL banners etc. go to stderr
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import torchvision

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = torchvision.models.resnet18(num_classes=10).to(dev).to(memory_format=torch.channels_last)
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=args.lr)
    crit = torch.nn.CrossEntropyLoss()
    wire_dtype = torch.bfloat16 if args.wire == "bf16" else torch.float32
    float_state = [v for v in model.state_dict().values() if v.is_floating_point()]
    int_state = [v for v in model.state_dict().values() if not v.is_floating_point()]
    n_float = sum(v.numel() for v in float_state)
    wire = torch.empty(n_float, dtype=wire_dtype, device=dev)
    counts = torch.zeros(world, device=dev)
    flat = None
    if args.graph:
        # one flat fp32 buffer behind every float state_dict entry (conv weights keep their channels_last strides)
        flat = torch.empty(n_float, dtype=torch.float32, device=dev)
        scaled = torch.empty_like(flat)
        off = 0
        with torch.no_grad():
            entries = [(n_, p_, True) for n_, p_ in model.named_parameters()] + \
                      [(n_, b_, False) for n_, b_ in model.named_buffers() if b_.is_floating_point()]
            order = {id(v): i for i, v in enumerate(float_state)}
            for name, t, is_param in entries:
                n_el = t.numel()
                v = flat[off: off + n_el]
                v = v.view(t.shape[0], t.shape[2], t.shape[3], t.shape[1]).permute(0, 3, 1, 2) if t.dim() == 4 else v.view(t.shape)
                v.copy_(t)
                if is_param:
                    t.data = v
                else:
                    owner = model
                    for part in name.split(".")[:-1]:
                        owner = getattr(owner, part)
                    owner._buffers[name.split(".")[-1]] = v
                off += n_el
        assert off == n_float
        float_state = [v for v in model.state_dict().values() if v.is_floating_point()]
        opt = torch.optim.SGD(model.parameters(), lr=args.lr)

    g = torch.Generator().manual_seed(1234 + rank)
    probs = torch._standard_gamma(torch.full((10,), 0.5), generator=g)
    probs = probs / probs.sum()
    y_host = torch.multinomial(probs, args.samples, replacement=True, generator=g).pin_memory()
    means = torch.randn(10, 3, 32, 32, generator=torch.Generator().manual_seed(7)) * 0.5
    X_host = (means[y_host] + torch.randn(args.samples, 3, 32, 32, generator=g)).to(torch.bfloat16).pin_memory()
    if args.graph:      # resident shard kept NHWC so a gathered batch is already channels_last (no per-step layout copy)
        X_host = X_host.permute(0, 2, 3, 1).contiguous().pin_memory()
    X_res, y_res = X_host.to(dev), y_host.to(dev)
    h2d = X_host.numel() * 2 + y_host.numel() * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def local_train(X, y, n_epoch):
        idxs = torch.randperm(X.shape[0], device=dev)
        hist = torch.zeros(n_epoch, device=dev)
        for epoch in range(n_epoch):
            nb = 0
            for b in torch.split(idxs, args.batch_size):
                opt.zero_grad(set_to_none=True)
                xb = X[b].contiguous(memory_format=torch.channels_last)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = model(xb)
                loss = crit(out.float(), y[b])
                hist[epoch] += loss.detach()
                loss.backward()
                opt.step()
                nb += 1
            hist[epoch] /= nb
        return hist

    graph_state = {}

    def graphed_train(X, y, n_epoch):
        """One captured graph per epoch: static shard / permutation buffers, the graph is replayed per epoch."""
        n_steps = X.shape[0] // args.batch_size
        gs = graph_state
        if not gs:
            gs["X"] = torch.empty_like(X)
            gs["y"] = torch.empty_like(y)
            gs["perm"] = torch.zeros(n_steps * args.batch_size, dtype=torch.int64, device=dev)
            gs["loss"] = torch.zeros((), device=dev)
            gs["X"].copy_(X); gs["y"].copy_(y)
            gs["perm"].copy_(torch.arange(n_steps * args.batch_size, device=dev) % X.shape[0])

            def one_epoch():
                for s_ in range(n_steps):
                    idx = gs["perm"][s_ * args.batch_size:(s_ + 1) * args.batch_size]
                    xb = gs["X"].index_select(0, idx).permute(0, 3, 1, 2)     # NHWC storage = channels_last view
                    yb = gs["y"].index_select(0, idx)
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        out = model(xb)
                    loss = crit(out.float(), yb)
                    gs["loss"] += loss.detach()
                    loss.backward()
                    opt.step()

            snap = flat.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                 # warm-up outside capture (cuDNN autotune, allocator)
                    gs["perm"][: args.batch_size] += 0
                    idx = gs["perm"][: args.batch_size]
                    xb = gs["X"].index_select(0, idx).permute(0, 3, 1, 2)
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        out = model(xb)
                    crit(out.float(), gs["y"].index_select(0, idx)).backward()
                    opt.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                one_epoch()
            gs["graph"] = g_
            flat.copy_(snap)
            for v in int_state:
                v.zero_()
        gs["X"].copy_(X, non_blocking=True)
        gs["y"].copy_(y, non_blocking=True)
        hist = torch.zeros(n_epoch, device=dev)
        for epoch in range(n_epoch):
            gs["perm"].copy_(torch.randperm(X.shape[0], device=dev)[: gs["perm"].numel()])
            gs["loss"].zero_()
            gs["graph"].replay()
            hist[epoch] = gs["loss"] / n_steps
        return hist

    @torch.no_grad()
    def aggregate_flat(n_k):
        counts.zero_()
        counts[rank] = n_k
        if world > 1:
            dist.all_reduce(counts)
        torch.mul(flat, counts[rank] / counts.sum(), out=scaled)
        wire.copy_(scaled)
        if world > 1:
            dist.all_reduce(wire)
        flat.copy_(wire)
        if world > 1 and int_state:
            ints = torch.stack([v.reshape(()) for v in int_state])
            dist.all_reduce(ints, op=dist.ReduceOp.MAX)
            for v, t_ in zip(int_state, ints):
                v.copy_(t_)

    @torch.no_grad()
    def aggregate(n_k):
        if flat is not None:
            return aggregate_flat(n_k)
        counts.zero_()
        counts[rank] = n_k
        if world > 1:
            dist.all_reduce(counts)
        w = counts[rank] / counts.sum()
        cat = torch.cat([v.reshape(-1).float() for v in float_state]).mul_(w)
        wire.copy_(cat)
        if world > 1:
            dist.all_reduce(wire)
        off = 0
        for v in float_state:
            v.copy_(wire[off: off + v.numel()].view_as(v))
            off += v.numel()
        if world > 1:
            for v in int_state:
                dist.all_reduce(v, op=dist.ReduceOp.MAX)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rounds(k, from_host, read):
        last = None
        for _ in range(k):
            flush.zero_()
            if from_host:
                X, y = X_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)
            else:
                X, y = X_res, y_res
            hist = (graphed_train if args.graph else local_train)(X, y, args.local_epochs)
            aggregate(float(args.samples))
            if read:
                last = hist.tolist()
        return last

    rounds(max(args.warmup, 3), False, False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rounds(args.steps, False, False)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    agg = []
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a0.record()
        aggregate(float(args.samples))
        a1.record()
        torch.cuda.synchronize()
        agg.append(a0.elapsed_time(a1))
    rounds(2, True, True)
    barrier()
    t0 = time.perf_counter()
    last = rounds(args.steps, True, True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([dev_ms, e2e_ms, min(agg) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, agg_us = [float(x) for x in t.tolist()]
    if rank == 0:
        per_round = world * args.samples * args.local_epochs
        os.write(real_stdout, (json.dumps({
            "impl": "baseline",
            "metric": "federated local samples/sec (whole box), ResNet-18 FedAvg, synthetic non-IID 32x32 shards",
            "value": per_round * args.steps / (dev_ms / 1e3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "rounds_per_s": args.steps / (dev_ms / 1e3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 (autocast)", "data": "synthetic",
            "config": {"model": "torchvision resnet18(num_classes=10)", "global_batch": world * args.batch_size,
                       "batch_size": args.batch_size, "samples_per_client": args.samples, "image": "32x32x3",
                       "local_epochs": args.local_epochs, "parallelism": "fedavg dp{}".format(world),
                       "backend": "nccl all_reduce + cuDNN/cuBLAS " + ("CUDA-graphed epoch, flat state buffer" if args.graph
                                                                       else "eager"),
                       "cuda_graph": bool(args.graph), "wire_dtype": args.wire,
                       "l2": "256 MiB memset between rounds (flush)"},
            "e2e": {"value": per_round * args.steps / (e2e_ms / 1e3), "unit": "samples/s",
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * args.local_epochs},
            "agg_bcast_us_per_round": agg_us, "final_loss": last[-1] if last else None,
        }) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
