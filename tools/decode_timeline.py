"""Kernel timeline of the CUDA-graph decode step (torch.profiler / CUPTI activity records): per-kernel average
durations inside graph replay (with programmatic dependent launch overlap) and the idle gaps between kernels.
    python tools/decode_timeline.py [tokens]
"""
import os
import sys
import collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    cfg = bench.MODEL_4B
    torch.cuda.set_device(0)
    from cogview_b200.generation import sampling
    model = bench.build_model(cfg, 1089, "cuda").eval()
    with torch.no_grad():
        sampling.filling_sequence(model, bench.make_template(4, 16, seed=0).cuda(), bench.SampleArgs)
        torch.cuda.synchronize()
        tmpl = bench.make_template(4, n, seed=0).cuda()
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            sampling.filling_sequence(model, tmpl, bench.SampleArgs)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "mem" not in e.name.lower()[:6]]
    evs.sort(key=lambda e: e.time_range.start)
    print("kernel records:", len(evs))
    # keep the steady-state decode part: records after the last prefill attention kernel
    last_prefill = max((i for i, e in enumerate(evs) if "attn_fwd_kernel" in e.name), default=0)
    dec = evs[last_prefill + 1:]
    t0, t1 = dec[0].time_range.start, dec[-1].time_range.end
    busy = collections.defaultdict(float)
    cnt = collections.Counter()
    def short(name):
        name = name.replace("void ", "").replace("(anonymous namespace)::", "")
        for cut in ("<", "("):
            if cut in name:
                name = name[:name.index(cut)]
        return name.split("::")[-1][:60] or "?"

    for e in dec:
        k = short(e.name)
        busy[k] += e.time_range.end - e.time_range.start
        cnt[k] += 1
    # union of busy intervals (kernels overlap under PDL) and gaps
    cover, cur_end, gaps = 0.0, dec[0].time_range.start, []
    for e in dec:
        s, en = e.time_range.start, e.time_range.end
        if s > cur_end:
            gaps.append(s - cur_end)
            cover += en - s
            cur_end = en
        elif en > cur_end:
            cover += en - cur_end
            cur_end = en
    steps = max(1, cnt.get("attn_decode_kernel", cfg["num_layers"]) // cfg["num_layers"])
    wall = t1 - t0
    print(f"decode window {wall / 1e3:.2f} ms, ~{steps} steps -> {wall / steps:.0f} us/step; "
          f"GPU covered {100 * cover / wall:.1f}%, {len(gaps)} gaps, total gap {sum(gaps) / 1e3:.2f} ms, "
          f"mean gap {sum(gaps) / max(1, len(gaps)):.2f} us")
    for k, v in sorted(busy.items(), key=lambda kv: -kv[1])[:22]:
        print(f"  {v / steps:9.1f} us/step  n/step={cnt[k] / steps:7.1f}  avg={v / cnt[k]:7.2f} us  {k}")


if __name__ == "__main__":
    main()
