"""CPU-side checks of round-2 host logic: the optimizer split point, the collective roofline arithmetic of bench.py and the
NVLink measurement's degenerate case."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sgd_split_point_follows_the_model_prefix(monkeypatch):
    from baton_b200.models import resnet18
    from baton_b200.parallel.arena import ParamArena
    from baton_b200.train import GraphedLocalSGD
    m = resnet18(10)
    arena = ParamArena(m, torch.device("cpu"))
    monkeypatch.setenv("BATON_SGD_OVERLAP", "1")
    tr = GraphedLocalSGD(m, arena, loss="ce", use_graph=False)
    stem = m.conv1.weight.numel() + m.bn1.weight.numel() + m.bn1.bias.numel()
    for prefix, want in (("layer1.", arena.slots["layer1.0.conv1.weight"].offset),
                         ("layer3.", arena.slots["layer3.0.conv1.weight"].offset)):
        m.tail_split_prefix = prefix
        tr._split = None
        split = tr._tail_split()
        assert split % 8 == 0 and want - 8 < split <= want
        if prefix == "layer1.":
            assert split >= stem - 8          # everything but the stem is in the overlapped slice
    monkeypatch.setenv("BATON_SGD_OVERLAP", "0")
    assert GraphedLocalSGD(m, arena, loss="ce", use_graph=False)._tail_split() == 0


def test_collective_roofline_uses_the_measured_link():
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    wire = 22_384_640
    r = bench._roofline(150.0, wire, 8, 673.0)
    assert "measured in this run" in r["bound"]
    floor = (7 / 8) * wire / 673e9 * 1e6
    assert abs(r["floor_us"] - floor) < 1e-6 and abs(r["fraction_of_measured"] - floor / 150.0) < 1e-9
    r = bench._roofline(150.0, wire, 8)                 # no measurement: the profiling guide's 770 GB/s
    assert "770" in r["bound"] and r["floor_us"] > 0
    r1 = bench._roofline(85.0, wire, 1)                  # one GPU: the bound is local HBM
    assert r1["bound"] == "hbm" and 0 < r1["fraction_of_measured"] < 1
    assert bench._roofline(0.0, wire, 8) is None


def test_link_measurement_needs_peers():
    from baton_b200.parallel.symm import SymmetricBuffer
    buf = SymmetricBuffer(1 << 12, "cpu")
    assert buf.world == 1 and buf.measure_link_gbps() is None and not buf.has_multicast
