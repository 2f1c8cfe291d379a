"""Times cv_attn_fwd / cv_attn_bwd at the 4B training shape (b=4, 40 heads, s=1088), with and without dropout."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cogview_b200 import ops


def timeit(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    b, heads, s = 4, 40, 1088
    h = heads * 64
    qkv = torch.randn((b, s, 3 * h), device="cuda").to(torch.bfloat16)
    d_out = torch.randn((b, s, h), device="cuda").to(torch.bfloat16)
    q, k, v = qkv[..., :h], qkv[..., h:2 * h], qkv[..., 2 * h:]
    for p in (0.0, 0.1):
        drop = (p, 1234, 3) if p > 0 else None
        r = ops.attn_fwd(q, k, v, heads, want_lse=True, dropout=drop)
        out, lse, mask = (r + (None,))[:3]
        tf = timeit(lambda: ops.attn_fwd(q, k, v, heads, want_lse=True, dropout=drop))
        tb = timeit(lambda: ops.attn_bwd(q, k, v, out, d_out, lse, heads, dropout_p=p, drop_mask=mask))
        print(f"dropout {p}: fwd {tf:.1f} us  bwd {tb:.1f} us (bwd includes delta + dq-store helpers)", flush=True)


if __name__ == "__main__":
    main()
