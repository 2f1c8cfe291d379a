import os, sys
sys.path.insert(0, "/root/repo")
import torch
from cogview_b200 import ops
M, N, K = 4352, 7680, 2560
x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
ws = [torch.randn((N, K), device="cuda").to(torch.bfloat16) * 0.02 for _ in range(4)]
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
def t(name, **kw):
    for w in ws[:2]: ops.gemm(x, w, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for w in ws: ops.gemm(x, w, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print("%-28s %8.1f us  %7.1f TFLOP/s" % (name, us, 2.0 * M * N * K / us / 1e6), flush=True)
t("plain"); t("bias", bias=bias)
