"""Phase anatomy of the persistent decode step (cv_decode_step): per-layer %globaltimer stamps of every CTA.

    python tools/step_prof.py [--layers 8] [--batch 4] [--t 512]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

NAMES = ["glueA", "qkv", "bar1", "attn", "bar2", "dense", "bar3", "glueB", "fc1", "bar4", "fc2+merge", "bar5"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--t", type=int, default=512)
    ap.add_argument("--hidden", type=int, default=2560)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from bench import build_model
    from cogview_b200 import ops
    from cogview_b200.mpu import kv_cache
    from cogview_b200.mpu.decode import DecodeRunner
    cfg = dict(num_layers=a.layers, vocab_size=58240, hidden_size=a.hidden, num_attention_heads=a.hidden // 64,
               max_sequence_length=1089)
    model = build_model(cfg, 1089, "cuda").eval()
    c = kv_cache._Caches(model.transformer, a.batch, torch.device("cuda"))
    c.buf.normal_()
    c.t = a.t
    r = DecodeRunner(model, c, use_graph=False)
    r._check_params()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    G = sms
    prof = torch.zeros((G, a.layers, 32), dtype=torch.int64, device="cuda")
    r.ids.fill_(5)
    r.pos.fill_(a.t)
    r.cur_len.fill_(a.t)

    def run(p):
        ops.decode_step(r.layer_table, len(r.params), r.heads, r.eps, r.fl[2], r.wte, r.wpe, r.fl[0], r.fl[1], r.ids,
                        r.pos, r.cur_len, c.buf, r.step_logits, r.workspace, prof=p)
    for _ in range(3):
        run(None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run(None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    wbytes = sum(p.numel() * 2 for p in model.parameters())
    kv = a.layers * a.batch * a.t * 2 * a.hidden * 2
    print("step %.1f us for %d layers + logits: %.1f MB weights + %.1f MB K|V -> %.0f GB/s" % (
        ms * 1e3, a.layers, wbytes / 1e6, kv / 1e6, (wbytes + kv) / ms / 1e6))
    run(prof)
    torch.cuda.synchronize()
    P = prof.cpu().double()
    d = (P[:, :, 1:13] - P[:, :, 0:12]) / 1e3           # us per phase [G, L, 12]
    lay = (P[:, :, 12] - P[:, :, 0]) / 1e3
    print("per-layer time (us): mean over CTAs per layer:", [round(x, 1) for x in lay.mean(0).tolist()])
    print("%-10s %8s %8s %8s" % ("phase", "mean", "min-cta", "max-cta"))
    for i, n in enumerate(NAMES):
        x = d[:, 1:, i]                                # skip layer 0 (cold start)
        print("%-10s %8.2f %8.2f %8.2f" % (n, x.mean().item(), x.mean(1).min().item(), x.mean(1).max().item()))
    print("sum of phase means: %.1f us/layer" % d[:, 1:, :].mean((0, 1)).sum().item())
    g = P[:, 1:, :]
    sub = [(g[:, :, 13] - g[:, :, 0]), (g[:, :, 14] - g[:, :, 13]), (g[:, :, 15] - g[:, :, 14]), (g[:, :, 1] - g[:, :, 15])]
    print("glueA anatomy (us): loads+reduce0 %.2f | dev+reduce1+v+reduce2 %.2f | dev2+reduce3 %.2f | xn+sync %.2f" % tuple(
        (x.mean().item() / 1e3) for x in sub))
    acct(P)


def acct(P):
    # wait accounting of the four linears: cycles thread 0 (stage group 0) / thread 256 (group 1) spent waiting for
    # ring data and in the per-tile barrier
    A = P[:, 1:, 16:32].mean((0, 1)) / 1.9e3          # us at ~1.9 GHz
    for gi, gname in enumerate(("group0", "group1")):
        print(gname, " ".join("%s: data-wait %.2f tile-bar %.2f |" % (n, A[gi * 8 + 2 * i].item(), A[gi * 8 + 2 * i + 1].item())
                              for i, n in enumerate(("qkv", "dense", "fc1", "fc2"))))


if __name__ == "__main__":
    main()
