"""One launch of each hot kernel at CogView-4B shapes, for `ncu --set full` captures."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogview_b200 import ops

torch.manual_seed(0)
h, M = 2560, 4
shapes = [("qkv", 3 * h, h), ("out", h, h), ("h4", 4 * h, h), ("h", h, 4 * h)]
ws = {n: (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16) for n, N, K in shapes}
xs = {K: torch.randn((M, K), device="cuda").to(torch.bfloat16) for K in (h, 4 * h)}
big = torch.randn((4352, h), device="cuda").to(torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "decode"
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    if which == "decode":
        for n, N, K in shapes:
            flush.zero_()
            ops.linear_small_m(xs[K], ws[n])
    else:
        for n, N, K in shapes[:3]:
            flush.zero_()
            ops.gemm(big, ws[n])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
