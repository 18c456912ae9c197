"""Per-node cost of small kernels inside a CUDA graph (with / without PDL: BATON_PDL=0|1), and
steady-state time of representative GEMM shapes.  All CUDA-event timed, warm, graph-replayed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.ops import functional as F  # noqa: E402

dev = torch.device("cuda:0")
BF16 = torch.bfloat16


def time_graph(fn, reps=200, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)   # us per call


print("PDL =", os.environ.get("BATON_PDL", "1"))
x = torch.randn(4096, device=dev)
xb = torch.empty(4096, device=dev, dtype=BF16)
print("tiny cast kernel      : {:6.2f} us/node".format(time_graph(lambda: F.cast(x, BF16, out=xb))))
t = torch.zeros(4096, device=dev)
print("torch add_ (native)   : {:6.2f} us/node".format(time_graph(lambda: t.add_(1.0))))
shapes = [("stem  fwd", 32768, 64, 152, False, False), ("l1 fwd", 8192, 64, 576, False, False),
          ("l2 fwd", 2048, 128, 1152, False, False), ("l3 fwd", 512, 256, 2304, False, False),
          ("l4 fwd", 128, 512, 4608, False, False), ("l4 dgrad", 128, 4608, 512, False, True),
          ("l1 dgrad", 8192, 576, 64, False, True), ("big", 8192, 8192, 8192, False, False),
          ("bert qkv", 4096, 2304, 768, False, False), ("bert ffn1", 4096, 3072, 768, False, False),
          ("bert ffn2", 4096, 768, 3072, False, False), ("bert dgrad", 4096, 768, 2304, False, True)]
for name, M, N, K, amn, bmn in shapes:
    A = torch.randn(M, K, device=dev).to(BF16)
    B = (torch.randn(K, N, device=dev) if bmn else torch.randn(N, K, device=dev)).to(BF16)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    us = time_graph(lambda: F.gemm(A, B, b_mn=bmn, out=out), reps=50 if M * N * K < 1e11 else 5)
    fl = 2.0 * M * N * K / (us * 1e-6) / 1e12
    print("gemm {:9s} {:6d}x{:5d}x{:5d}: {:8.2f} us  {:8.1f} TFLOP/s".format(name, M, N, K, us, fl))
# wgrad (split-K accumulate fp32)
for name, M, N, K in [("l1 wgrad", 64, 576, 8192), ("l4 wgrad", 512, 4608, 128)]:
    A = torch.randn(K, M, device=dev).to(BF16)
    B = torch.randn(K, N, device=dev).to(BF16)
    out = torch.zeros(M, N, device=dev)
    us = time_graph(lambda: F.gemm(A, B, a_mn=True, b_mn=True, out=out, accumulate=True), reps=50)
    print("gemm {:9s} {:6d}x{:5d}x{:5d}: {:8.2f} us".format(name, M, N, K, us))
# cuBLAS reference for the same shapes
for name, M, N, K, amn, bmn in shapes:
    A = torch.randn(M, K, device=dev).to(BF16)
    B = torch.randn(N, K, device=dev).to(BF16)
    out = torch.empty(M, N, device=dev, dtype=BF16)
    us = time_graph(lambda: torch.matmul(A, B.t(), out=out), reps=50 if M * N * K < 1e11 else 5)
    print("cublas {:9s}: {:8.2f} us  {:8.1f} TFLOP/s".format(name, us, 2.0 * M * N * K / (us * 1e-6) / 1e12))
