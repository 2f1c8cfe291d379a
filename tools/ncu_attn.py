"""One attention forward + backward at the CogView-4B layer shape for `ncu --set full --import-source on`."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogview_b200 import ops
b, s, heads = 4, 1088, 40
h = heads * 64
qkv = (torch.randn((b, s, 3 * h), device="cuda") * 0.5).to(torch.bfloat16)
d_out = torch.randn((b, s, h), device="cuda").to(torch.bfloat16)
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    out, lse = ops.attn_fwd(qkv[..., :h], qkv[..., h:2 * h], qkv[..., 2 * h:], heads, want_lse=True)
    ops.attn_bwd(qkv[..., :h], qkv[..., h:2 * h], qkv[..., 2 * h:], out, d_out, lse, heads)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# hot timings
for name, fn in (("fwd", lambda: ops.attn_fwd(qkv[..., :h], qkv[..., h:2 * h], qkv[..., 2 * h:], heads, want_lse=True)),
                 ("bwd", lambda: ops.attn_bwd(qkv[..., :h], qkv[..., h:2 * h], qkv[..., 2 * h:], out, d_out, lse, heads))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.1f us" % (e0.elapsed_time(e1) * 100))
