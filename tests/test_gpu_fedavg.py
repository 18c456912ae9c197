"""Distributed tier: the fused NVLink FedAvg collective at 2..N GPUs (torchrun worker in
tests/mp_fedavg_check.py) and its single-GPU degenerate case."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BF16 = torch.bfloat16


@pytest.mark.gpu
def test_fedavg_kernel_world1_is_identity_and_applies_delta():
    from baton_b200.models import MLP2
    from baton_b200.parallel.arena import ParamArena
    from baton_b200.parallel.fedavg import FedAvgSession
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for wire, mode in (("bf16", "delta"), ("fp32", "weights"), ("bf16", "weights"), ("fp32", "delta")):
        m = MLP2(64, 256, 8)
        arena = ParamArena(m, dev, momentum=True)
        sess = FedAvgSession(arena, wire_dtype=wire, mode=mode, n_ctas=8)
        g0 = arena.global_w.clone()
        arena.theta.add_(torch.randn_like(arena.theta) * 0.01)
        want = arena.theta.clone()
        arena.momentum.fill_(3.0)
        sess.aggregate(my_n=128.0)
        torch.cuda.synchronize()
        sess.check()
        tol = 1e-6 if wire == "fp32" else (3e-4 if mode == "delta" else 8e-3)   # bf16: |delta| * 2^-9
        assert float((arena.theta - want).abs().max()) < tol, (wire, mode)
        assert torch.equal(arena.theta, arena.global_w)
        assert torch.equal(arena.theta_bf16, arena.theta.to(BF16))
        assert float(arena.momentum.abs().max()) == 0.0
        # n_k = 0 on the only rank: nothing to average, the replica keeps the global model
        sess.aggregate(n_samples_by_rank=[0.0])
        torch.cuda.synchronize()
        # state_dict views still alias the arena
        assert m.fc1.weight.data_ptr() == arena.theta.data_ptr() + arena.slots["fc1.weight"].offset * 4


@pytest.mark.gpu
def test_fedavg_fp8_block_scaled_wire_world1():
    """MXFP8 wire: the delta crosses the wire as e4m3 with one power-of-two scale per 32 elements (upload
    and broadcast are both quantised), so the applied update is within two e4m3 roundings of the true one."""
    from baton_b200.models import MLP2
    from baton_b200.parallel.arena import ParamArena
    from baton_b200.parallel.fedavg import FedAvgSession
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = MLP2(72, 250, 6)                       # arena size not a multiple of 32 somewhere inside
    arena = ParamArena(m, dev, momentum=True)
    sess = FedAvgSession(arena, wire_dtype="fp8", mode="delta", n_ctas=8)
    assert not sess.use_nvls and sess.wire_bytes() == arena.n + (arena.n + 31) // 32
    g0 = arena.global_w.clone()
    # per-block dynamic range: scale the drift differently per 32-element block
    drift = torch.randn_like(arena.theta) * 0.01
    blk = torch.arange(arena.n, device=dev) // 32
    drift *= torch.pow(2.0, (blk % 13).float() - 6.0)
    arena.theta.add_(drift)
    want = arena.theta.clone()
    sess.aggregate(my_n=64.0)
    torch.cuda.synchronize()
    sess.check()
    got_delta = arena.theta - g0
    true_delta = want - g0
    pad = (-arena.n) % 32
    td = torch.nn.functional.pad(true_delta, (0, pad)).view(-1, 32)
    gd = torch.nn.functional.pad(got_delta, (0, pad)).view(-1, 32)
    amax = td.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
    rel = ((gd - td).abs() / amax).max()
    assert float(rel) < 0.14, float(rel)              # 2 roundings x 2^-4, relative to the block maximum
    rms = float((gd - td).pow(2).mean().sqrt() / td.pow(2).mean().sqrt())
    assert rms < 0.06, rms
    assert torch.equal(arena.theta, arena.global_w)
    assert torch.equal(arena.theta_bf16, arena.theta.to(BF16))
    assert float(arena.momentum.abs().max()) == 0.0


@pytest.mark.gpu
@pytest.mark.multigpu
def test_fused_fedavg_multi_gpu_matches_formula_and_nccl_oracle():
    n = min(torch.cuda.device_count(), 8)
    port = 29500 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_fedavg_check.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    tail = "\n".join(proc.stdout.splitlines()[-60:])
    assert proc.returncode == 0 and "RESULT PASS" in proc.stdout, tail


@pytest.mark.gpu
@pytest.mark.multigpu
def test_manager_and_gpu_workers_over_http_survive_a_dead_seat():
    """Baton's API on the NVLink data plane at >= 2 seats (tests/mp_api_check.py): CPU manager + one GpuExperimentWorker
    per GPU over HTTP; initial model distribution, a killed seat (eviction, survivors aggregate, no hang) and its return."""
    n = min(torch.cuda.device_count(), 4)
    port = 29500 + ((os.getpid() + 317) % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_api_check.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    tail = "\n".join(proc.stdout.splitlines()[-60:])
    assert proc.returncode == 0 and "RESULT PASS" in proc.stdout, tail


@pytest.mark.gpu
@pytest.mark.parametrize("wire,mode", [("bf16", "delta"), ("fp32", "delta"), ("bf16", "weights")])
def test_upload_copy_emitted_by_the_optimizer_matches_in_kernel_pack(wire, mode):
    """the last SGD step writes the client's wire copy itself (``fused_sgd(pack=...)``) and the collective
    skips its pack phase.  Same start state, same gradient: the round result must be bit-identical to the in-kernel
    pack, over several rounds (the wire is double-buffered by round parity, the address travels in a device word)."""
    from baton_b200.models import resnet18
    from baton_b200.ops import functional as F
    from baton_b200.parallel.arena import ParamArena
    from baton_b200.parallel.fedavg import FedAvgSession
    dev = torch.device("cuda:0")
    results = []
    for prepack in (False, True):
        torch.manual_seed(0)
        m = resnet18(10)
        arena = ParamArena(m, dev)
        sess = FedAvgSession(arena, wire_dtype=wire, mode=mode, n_ctas=32)
        hyper = torch.tensor([0.1, 0.0, 0.0, 0.0], device=dev)
        gen = torch.Generator(device=dev).manual_seed(7)
        for rnd in range(3):
            arena.grad.copy_(torch.randn(arena.n_param, device=dev, generator=gen) * 0.01)
            arena.theta[arena.n_param:].add_(0.001 * (rnd + 1))          # float buffers drift too (BatchNorm statistics)
            if prepack:
                sess.arm_prepack(64.0)
            F.fused_sgd(arena.theta[: arena.n_param], arena.grad, hyper, None, arena.theta_bf16[: arena.n_param],
                        pack=sess.pack_spec() if prepack else None)
            sess.aggregate(my_n=64.0, prepacked=prepack)
            assert sess.last_prepacked == prepack
        torch.cuda.synchronize()
        sess.check()
        results.append((arena.theta.clone(), arena.global_w.clone(), arena.theta_bf16.clone()))
    for a, b in zip(*results):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_bcast_gemm_inside_the_captured_epoch_resnet18():
    """K3 on the flagship path: with ``tile_flags=True`` the engine gates the first convolution (weight staging + TMA
    producer of its GEMM) on the collective's arrival flags through a device-resident epoch word, captures the epoch
    as two graphs and replays the first one while the collective is still on its side stream.  Training must behave
    exactly like the plain engine (same losses within the run-to-run noise of the atomics) and learn."""
    from baton_b200.data import ShardSpec, image_shard
    from baton_b200.models import resnet18
    from baton_b200.parallel.engine import FederatedEngine
    dev = torch.device("cuda:0")
    X, y = image_shard(ShardSpec(0, torch.full((10,), 0.1), 1024), noise=0.3)
    X, y = X.to(dev).to(BF16), y.to(dev)
    losses = {}
    for k3 in (False, True):
        torch.manual_seed(0)
        eng = FederatedEngine(resnet18(10), dev, backend="fused", lr=0.05, batch_size=128, tile_flags=k3, n_ctas=64)
        assert eng.k3 == k3
        hist = []
        for _ in range(6):
            hist += eng.run_round((X, y), n_epoch=1).loss_history
        eng.sync()
        torch.cuda.synchronize()
        eng.session.check()
        if k3:
            ent = next(iter(eng.trainer._graphs.values()))
            assert ent["graph2"] is not None                      # the epoch really is two graphs
            assert int(eng.session.tile_flags.min()) == eng.session.rounds
            assert eng.session.last_prepacked
        losses[k3] = hist
        assert torch.equal(eng.arena.theta, eng.arena.global_w)
    assert losses[True][-1] < losses[True][0] * 0.8, losses[True]
    assert abs(losses[True][0] - losses[False][0]) < 0.05 * abs(losses[False][0]) + 1e-3, losses
