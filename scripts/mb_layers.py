"""Per-layer forward cost of the ResNet-18 conv shapes as they run inside a captured step: each rep is
conv (implicit GEMM, statistics fused or not) followed by a full-GPU elementwise kernel (bn_apply), so consecutive
GEMMs cannot overlap each other the way identical back-to-back launches do.  Sweeps tile width / cluster split-K."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.ops import functional as F  # noqa: E402
from baton_b200.ops import load  # noqa: E402

dev = torch.device("cuda:0")
BF16 = torch.bfloat16
C = load()
REPS = 20


def time_graph(fn, iters=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def bench(name, n, h, cin, cout, k, stride, pad):
    x = torch.randn(n, h, h, cin, device=dev).to(BF16)
    w = (torch.randn(cout, k * k * cin, device=dev) * 0.05).to(BF16)
    ho = F.conv_out_size(h, k, stride, pad)
    M = n * ho * ho
    y = torch.empty(M, cout, device=dev, dtype=BF16)
    z = torch.empty(M, cout, device=dev, dtype=BF16)
    ws = torch.zeros(REPS, 4 * cout, device=dev)
    gamma, beta = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    rm, rv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
    sm, sr = torch.empty(cout, device=dev), torch.empty(cout, device=dev)

    def bn_only():
        for i in range(REPS):
            C.bn_apply(y, None, z, ws[i][: 2 * cout], gamma, beta, rm, rv, sm, sr, None, M, cout, 1e-5, 0.1, True, True)
    t_bn = time_graph(bn_only) / REPS
    out = ["{:10s} M={:5d} N={:3d} K={:4d}  bn_apply alone {:5.2f} us |".format(name, M, cout, k * k * cin, t_bn)]
    for bn in (64, 128, 256):
        if bn > 64 and cout < bn:
            continue
        for ck in (1, 2, 4, 8):
            for stats in (False, True):
                def run():
                    ws.zero_()
                    for i in range(REPS):
                        ok = C.conv_igemm_fwd(x, w, y, k, k, stride, pad, ho, ho, ck, bn, ws[i][: 2 * cout] if stats else None)
                        assert ok
                        C.bn_apply(y, None, z, ws[i][: 2 * cout], gamma, beta, rm, rv, sm, sr, None, M, cout, 1e-5, 0.1,
                                   True, True)
                try:
                    t = time_graph(run) / REPS - t_bn
                    out.append(" bn{} ck{}{} {:5.2f}".format(bn, ck, "+st" if stats else "   ", t))
                except Exception as e:  # unsupported combination
                    out.append(" bn{} ck{} ERR".format(bn, ck))
                    torch.cuda.synchronize()
    print("".join(out), flush=True)


shapes = [("l1", 128, 8, 64, 64, 3, 1, 1), ("l2.0c1", 128, 8, 64, 128, 3, 2, 1), ("l2", 128, 4, 128, 128, 3, 1, 1),
          ("l2.ds", 128, 8, 64, 128, 1, 2, 0), ("l3.0c1", 128, 4, 128, 256, 3, 2, 1), ("l3", 128, 2, 256, 256, 3, 1, 1),
          ("l4.0c1", 128, 2, 256, 512, 3, 2, 1)]
for s in shapes:
    bench(*s)
