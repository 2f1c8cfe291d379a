"""Bring-up helper: attention backward on a few shapes, each in a subprocess with a hard timeout."""
import subprocess
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [(1, 1, 128, 0), (1, 1, 256, 0), (1, 1, 200, 0), (1, 3, 128, 0), (1, 3, 200, 0), (2, 2, 384, 0), (1, 2, 300, 130)]


def run(i):
    import torch
    from cogview_b200 import ops
    from oracle import cogview_oracle as O
    b, heads, s, sep = CASES[i]
    g = torch.Generator().manual_seed(s + sep)
    h = heads * 64
    qkv = torch.randn((b, s, 3 * h), generator=g).to(torch.bfloat16)
    d_out = torch.randn((b, s, h), generator=g).to(torch.bfloat16)
    hv = lambda t: t.view(b, s, heads, 64).permute(0, 2, 1, 3)
    qr, kr, vr = (qkv[..., j * h:(j + 1) * h].float().clone().requires_grad_(True) for j in range(3))
    ref = O.standard_attention(hv(qr), hv(kr), hv(vr), O.build_sep_mask(s, s, sep)).permute(0, 2, 1, 3).reshape(b, s, h)
    ref.backward(d_out.float())
    qc = qkv.cuda()
    out, lse = ops.attn_fwd(qc[..., :h], qc[..., h:2 * h], qc[..., 2 * h:], heads, sep=sep, want_lse=True)
    torch.cuda.synchronize()
    dqkv = ops.attn_bwd(qc[..., :h], qc[..., h:2 * h], qc[..., 2 * h:], out, d_out.cuda(), lse, heads, sep=sep)
    torch.cuda.synchronize()
    errs = []
    for got, want in ((dqkv[..., :h], qr.grad), (dqkv[..., h:2 * h], kr.grad), (dqkv[..., 2 * h:], vr.grad)):
        errs.append(((got.float().cpu() - want).abs().max() / want.abs().max()).item())
    print("CASE", CASES[i], "rel errs dq dk dv", errs, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=90)
                print(r.stdout[-1500:], r.stderr[-800:] if r.returncode else "", flush=True)
            except subprocess.TimeoutExpired as e:
                print("CASE", CASES[i], "TIMEOUT", (e.stdout or b"")[-1500:], flush=True)
