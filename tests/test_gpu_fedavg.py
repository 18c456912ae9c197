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
