import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogview_b200 import ops
M, N, K = 4352, 10240, 2560
x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
ws = [torch.randn((N, K), device="cuda").to(torch.bfloat16) * 0.02 for _ in range(4)]
bias = torch.randn(N, device="cuda").to(torch.bfloat16)
am = torch.zeros(1, device="cuda")
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
t("plain")
t("bias", bias=bias)
t("bias+absmax", bias=bias, absmax=am)
t("bias+gelu", bias=bias, act=1)
t("bias+relu", bias=bias, act=2)
t("bias+preact", bias=bias, want_preact=True)
t("bias+gelu+preact", bias=bias, act=1, want_preact=True)
t("f32 out", out_dtype=torch.float32)
