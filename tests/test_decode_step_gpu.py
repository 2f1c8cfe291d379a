"""GPU tests of the persistent decode step (cv_decode_step, csrc/decode_step.cu) and the fused sampling epilogue
(cv_sample_topk, csrc/sample.cu).

* the one-kernel step against the per-operation decode path (cv_linear_small_m / cv_attn_decode /
  cv_ln_pair_small_m, each parity-tested against the oracle in test_kernels_gpu.py) on the same weights and cache,
  for the tiny config (h = 256: one K chunk per warp) and the 4B layer shape (h = 2560: five chunks), several
  batch sizes and memory lengths — logits within 1e-2 of the logit scale (both are bf16 paths that differ only
  in summation order), the appended K|V rows within bf16 rounding;
* the step against the fp32 oracle with hidden-state `mems` through GPT2Model (test_model_gpu.py::test_decode_with_mems
  runs through this kernel as well, since it is the default path for batch <= 8);
* the sampling kernel against the reference semantics (generation/sampling.py:24-33,157-183) restated with torch:
  same kept set (ties at the k-th value included), same probabilities, draws distributed accordingly.
"""
import math

import pytest
import torch

from oracle import recipes

pytestmark = pytest.mark.gpu


def _model(cfg, maxlen):
    from cogview_b200.model import GPT2Model
    m = GPT2Model(num_layers=cfg["num_layers"], vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
                  num_attention_heads=cfg["num_attention_heads"], embedding_dropout_prob=0.0,
                  attention_dropout_prob=0.0, output_dropout_prob=0.0, max_sequence_length=maxlen,
                  max_memory_length=maxlen, checkpoint_activations=False)
    sd = recipes.gpt2_state_dict(num_layers=cfg["num_layers"], vocab_size=cfg["vocab_size"],
                                 hidden_size=cfg["hidden_size"], max_sequence_length=maxlen, seed=7)
    m.load_state_dict(sd)
    return m.cuda().bfloat16().eval()


def _runners(model, b, t, monkeypatch):
    """Two DecodeRunners (persistent / per-op) on identical caches holding t random tokens."""
    from cogview_b200.mpu import kv_cache
    from cogview_b200.mpu.decode import DecodeRunner
    tr = model.transformer
    g = torch.Generator(device="cuda").manual_seed(100 + b + t)
    out = []
    base = None
    for persistent in (True, False):
        monkeypatch.setenv("COGVIEW_B200_PERSISTENT", "1" if persistent else "0")
        c = kv_cache._Caches(tr, b, torch.device("cuda"))
        if base is None:
            base = (torch.randn(c.buf.shape, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
        c.buf.copy_(base)
        c.t = t
        r = DecodeRunner(model, c, use_graph=False)
        assert r.persistent == persistent
        out.append((r, c))
    return out


CFG_TINY = dict(num_layers=2, vocab_size=58240, hidden_size=256, num_attention_heads=4)
CFG_WIDE = dict(num_layers=2, vocab_size=8192 + 512, hidden_size=2560, num_attention_heads=40)
CFG_MID = dict(num_layers=3, vocab_size=4096, hidden_size=1024, num_attention_heads=16)
CFG_512 = dict(num_layers=2, vocab_size=1000, hidden_size=512, num_attention_heads=8)
CFG_768 = dict(num_layers=2, vocab_size=1000, hidden_size=768, num_attention_heads=12)


@pytest.mark.parametrize("cfg,b,t", [(CFG_TINY, 1, 0), (CFG_TINY, 3, 17), (CFG_TINY, 4, 120), (CFG_TINY, 8, 64),
                                     (CFG_MID, 2, 33), (CFG_MID, 5, 250), (CFG_512, 4, 40), (CFG_768, 6, 77),
                                     (CFG_WIDE, 4, 0), (CFG_WIDE, 4, 300), (CFG_WIDE, 7, 129), (CFG_WIDE, 8, 1085)])
def test_persistent_step_matches_per_op_path(cfg, b, t, monkeypatch):
    maxlen = 1089 if t > 200 else 256
    model = _model(cfg, maxlen)
    (rp, cp), (ro, co) = _runners(model, b, t, monkeypatch)
    g = torch.Generator().manual_seed(t)
    ids = torch.randint(0, cfg["vocab_size"], (b, 1), generator=g).cuda()
    pos = torch.full((b, 1), t, dtype=torch.long, device="cuda")
    with torch.no_grad():
        for step in range(3):                       # three consecutive steps: counters / buffers are reusable
            lp = rp.step(ids, pos + step, t + step).clone()
            lo = ro.step(ids, pos + step, t + step).clone()
            torch.cuda.synchronize()
            scale = lo.abs().max().item()
            err = (lp - lo).abs().max().item()
            print("h=%d b=%d t=%d step %d: logits max|diff| %.3e (scale %.3e)" % (cfg["hidden_size"], b, t + step, step,
                                                                                 err, scale))
            assert bool(torch.isfinite(lp).all())
            assert err < 1e-2 * scale
            # appended K|V of the new token in every layer
            kp = cp.buf[:, :, t + step].float()
            ko = co.buf[:, :, t + step].float()
            assert (kp - ko).abs().max().item() < 2e-2 * ko.abs().max().item()
            # older cache rows are untouched
            if t > 0:
                assert torch.equal(cp.buf[:, :, :t], co.buf[:, :, :t])
            ids = lo[:, :recipes.IMG_VOCAB].argmax(-1, keepdim=True)


def test_persistent_step_graph_replay_and_weight_refresh(monkeypatch):
    """The step under CUDA-graph capture (the production mode) reproduces the eager launch, and a weight change
    after the first generation is picked up (pointer table + captured graph are rebuilt)."""
    from cogview_b200.mpu import kv_cache
    from cogview_b200.mpu.decode import DecodeRunner
    model = _model(CFG_TINY, 128)
    tr = model.transformer
    b, t = 4, 9
    base = None
    outs = []
    for use_graph in (False, True):
        c = kv_cache._Caches(tr, b, torch.device("cuda"))
        if base is None:
            base = torch.randn(c.buf.shape, device="cuda").to(torch.bfloat16)
        c.buf.copy_(base)
        c.t = t
        r = DecodeRunner(model, c, use_graph=use_graph)
        ids = torch.arange(b, device="cuda").view(b, 1) + 5
        pos = torch.full((b, 1), t, dtype=torch.long, device="cuda")
        with torch.no_grad():
            a = r.step(ids, pos, t).clone()
            bb = r.step(ids + 1, pos + 1, t + 1).clone()
        outs.append((a, bb, r, c))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # change a weight in place: the next sequence must see it
    a, _, r, c = outs[1]
    with torch.no_grad():
        model.transformer.final_layernorm.weight.mul_(2.0)
        ids = torch.arange(b, device="cuda").view(b, 1) + 5
        pos = torch.full((b, 1), t, dtype=torch.long, device="cuda")
        c.buf.copy_(base)
        a2 = r.step(ids, pos, t).clone()          # t <= last_t: parameters are re-checked
    assert not torch.allclose(a, a2)


def _ref_probs(logits, temperature, top_k, valid):
    lg = logits.clone() / temperature
    mask = torch.ones_like(lg, dtype=torch.bool)
    for lo, hi in valid:
        mask[:, lo:hi] = False
    lg[mask] = -float("inf")
    if top_k > 0:
        kth = torch.topk(lg, top_k)[0][..., -1, None]
        lg[lg < kth] = -float("inf")
    return torch.softmax(lg, dim=-1)


@pytest.mark.parametrize("valid,top_k,temperature", [([(0, 8192)], 200, 1.0), ([(8192, 58192)], 200, 0.8),
                                                     ([(0, 8192), (58192, 58240)], 50, 1.3), ([(0, 58240)], 0, 1.0),
                                                     ([(0, 8192)], 1, 1.0), ([(0, 8192)], 9000, 1.0)])
def test_sample_topk_distribution_matches_reference_semantics(valid, top_k, temperature):
    from cogview_b200 import ops
    torch.manual_seed(0)
    b, V = 4, 58240
    logits = (torch.randn((b, V), device="cuda") * 3.0).contiguous()
    # ties exactly at the top-k threshold: `logits < kth` keeps all of them
    lo0 = valid[0][0]
    srt = torch.sort(logits[0, valid[0][0]:valid[0][1]], descending=True)[0]
    if 1 < top_k < 500:
        logits[0, lo0 + 3] = srt[top_k - 1]
        logits[0, lo0 + 11] = srt[top_k - 1]
    keep = logits.clone()
    ref = _ref_probs(logits, temperature, top_k, valid)
    nxt, probs = ops.sample_topk(logits, temperature, top_k, valid, seed=123, want_probs=True)
    torch.cuda.synchronize()
    assert torch.equal(logits, keep), "the kernel must not modify the logits"
    assert torch.equal(probs > 0, ref > 0), "kept set differs from top_k_logits"
    assert (probs - ref).abs().max().item() < 1e-6 + 1e-4 * ref.max().item()
    assert bool(((nxt >= 0) & (nxt < V)).all()) and bool((ref.gather(1, nxt.view(-1, 1)) > 0).all())
    # draws follow the distribution: 4000 draws per row through the device-side step counter
    n = 4000
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.zeros((b, n), dtype=torch.int64, device="cuda")
    score = torch.zeros(b, device="cuda")
    pos = torch.zeros((b, 1), dtype=torch.int64, device="cuda")
    cur = torch.zeros(1, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    ids = torch.zeros(b, dtype=torch.int64, device="cuda")
    seed_dev = torch.tensor([991], dtype=torch.int64, device="cuda")
    for _ in range(n):
        ops.sample_topk(logits, temperature, top_k, valid, seed_dev=seed_dev, step=step, next_ids=ids, out_tokens=out,
                        score_acc=score, pos=pos, cur_len=cur, done=done)
    torch.cuda.synchronize()
    assert int(step) == n and int(cur) == n and bool((pos == n).all()) and int(done) == 0
    assert torch.equal(out[:, -1], ids)
    lp = torch.log(ref.gather(1, out)).sum(1)
    assert torch.allclose(score, lp, rtol=1e-3, atol=1e-2)
    for r in range(b):
        p = ref[r]
        top = torch.topk(p, 5)
        cnt = torch.bincount(out[r], minlength=V).float()
        assert float(cnt[p == 0].sum()) == 0
        for pv, iv in zip(top.values.tolist(), top.indices.tolist()):
            sd = math.sqrt(n * pv * (1 - pv)) + 1.0
            assert abs(float(cnt[iv]) - n * pv) < 5 * sd, (r, iv, pv, float(cnt[iv]))
