"""Profiling driver: a few 4B training steps or decode steps between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none ... python tools/profile_step.py train
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "train"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dropout = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    cfg = bench.MODEL_4B
    torch.cuda.set_device(0)
    if what == "train":
        from cogview_b200 import mpu
        from cogview_b200.model import gpt2_get_params_for_weight_decay_optimization
        from cogview_b200.optim import FusedAdamW
        model = bench.build_model(cfg, 0, "cuda", dropout=dropout).train()
        groups = gpt2_get_params_for_weight_decay_optimization(model)
        opt = FusedAdamW(groups, lr=4e-4, weight_decay=0.01, max_grad_norm=1.0)
        b, s = 4, 1088
        tok = torch.randint(0, 8192, (b, s + 1), device="cuda")
        pos = torch.arange(s, device="cuda").unsqueeze(0).expand(b, -1).contiguous()
        mask = torch.tril(torch.ones((1, 1, s, s), device="cuda"))

        def step():
            logits, *_ = model(tok[:, :-1].contiguous(), pos, mask, None, None, 0)
            loss = mpu.vocab_parallel_cross_entropy(logits, tok[:, 1:].contiguous()).mean()
            for p in model.parameters():
                p.grad = None
            loss.backward()
            opt.step()
        step()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        os.environ["COGVIEW_B200_CUDA_GRAPH"] = "0"     # ncu sees individual launches
        from cogview_b200.generation import sampling
        model = bench.build_model(cfg, 1089, "cuda").eval()
        tmpl = bench.make_template(4, 24, seed=0).cuda()
        with torch.no_grad():
            sampling.filling_sequence(model, tmpl, bench.SampleArgs)
            torch.cuda.synchronize()
            tmpl2 = bench.make_template(4, 8 + n, seed=0).cuda()
            torch.cuda.profiler.start()
            sampling.filling_sequence(model, tmpl2, bench.SampleArgs)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
