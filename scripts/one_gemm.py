"""Launch one bf16 GEMM shape a few times (for `ncu -k regex:gemm_bf16 -s 2 -c 1 --set full ...`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baton_b200.ops import functional as F  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
b = torch.randn(N, K, device=dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d = F.gemm(a, b)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("gemm {}x{}x{}: best {:.1f} us  {:.1f} TFLOP/s".format(M, N, K, min(ts), 2.0 * M * N * K / min(ts) / 1e6))
