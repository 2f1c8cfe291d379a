"""GPU parity of sparse TRAINING attention (sparse_attention + _chunk,
mpu/sparse_transformer.py:675-725, :629-650, pivot mask :491-496 / :569) — cv_attn_sparse_fwd / cv_attn_sparse_bwd.

* forward against tests/golden/attention.npz:sparse_train, the UNMODIFIED reference's output (s = 512, w = 64,
  times = 3, 96 pivots, 2 x 3 heads), and against the oracle restatement on bf16-rounded inputs;
* backward (dq, dk, dv) against the oracle's autograd through O.sparse_attention with the reference's own
  rmask-gathered pivot mask, at the golden shape and at the CogView-sr shape (s = 4096, w = 128, times = 6, 768 pivots);
* a 2-layer GPT2Model training step with is_sparse = 1 (fresh pivots per checkpointed layer in the reference's
  random.sample order) against the oracle evaluated with the same pivots.
Tolerances (bf16 tensor-core math vs the fp32 oracle): 2e-2 of the output scale forward, 4e-2 of each gradient's scale."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import cogview_oracle as O
from oracle import recipes

pytestmark = pytest.mark.gpu


def _rmask(s, w, times):
    g = s // w
    tmp = torch.ones((g - times + 1, w, w))
    tmp = torch.tril(1 - torch.block_diag(*tmp))
    return torch.nn.functional.pad(tmp, (0, (times - 1) * w, (times - 1) * w, 0))


def _pivots(b, s, n_piv, txt_n, seed):
    random.seed(seed)
    return torch.stack([
        torch.cat((torch.arange(0, txt_n[i]),
                   torch.tensor(random.sample(range(txt_n[i], s), n_piv - txt_n[i]), dtype=torch.long)))
        for i in range(b)])


def test_sparse_forward_matches_reference_golden(golden_dir):
    from cogview_b200.mpu.sparse_transformer import sparse_attention
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    b, nh, s, hn, w, times, n_piv, _ = [int(x) for x in g["dims"]]
    gen = torch.Generator().manual_seed(int(g["seed"]))
    q, k, v = (torch.randn((b, nh, s, hn), generator=gen) for _ in range(3))
    pivot_idx = torch.from_numpy(g["pivot_idx"])
    out = sparse_attention(q.cuda().bfloat16(), k.cuda().bfloat16(), v.cuda().bfloat16(), pivot_idx.cuda(), None, w,
                           times).float().cpu()
    ref = g["sparse_train"]
    scale = np.abs(ref).max()
    err = np.abs(out.numpy()[:, :, ::7] - ref).max()
    print("sparse fwd vs reference golden: max|diff| %.3e (scale %.3e)" % (err, scale))
    assert err < 2e-2 * scale
    # and against the oracle on the same bf16-rounded inputs (removes the input rounding from the comparison)
    qb, kb, vb = (t.bfloat16().float() for t in (q, k, v))
    pam = _rmask(s, w, times).expand(b, s, s).gather(dim=-1, index=pivot_idx.unsqueeze(1).expand(b, s, n_piv))
    o = O.sparse_attention(qb, kb, vb, pivot_idx, pam, w, times)
    assert (out - o).abs().max().item() < 1e-2 * o.abs().max().item()


@pytest.mark.parametrize("b,nh,s,w,times,n_piv,txt", [(2, 3, 512, 64, 3, 96, (48, 20)), (1, 2, 1024, 128, 6, 200, (64,)),
                                                     (1, 2, 4096, 128, 6, 768, (100,)), (2, 1, 256, 128, 2, 40, (0, 7)),
                                                     (1, 1, 768, 128, 6, 64, (30,))])
def test_sparse_forward_backward_match_oracle_autograd(b, nh, s, w, times, n_piv, txt):
    from cogview_b200 import ops
    gen = torch.Generator().manual_seed(s + n_piv)
    h = nh * 64
    qkv = torch.randn((b, s, 3 * h), generator=gen).bfloat16()
    d_out = (torch.randn((b, s, h), generator=gen) * 0.5).bfloat16()
    pivot_idx = _pivots(b, s, n_piv, txt, seed=99)
    qc = qkv.cuda()
    ctx, lse = ops.attn_sparse_fwd(qc[..., :h], qc[..., h:2 * h], qc[..., 2 * h:], nh, pivot_idx.cuda(), w, times,
                                   want_lse=True)
    dqkv = ops.attn_sparse_bwd(qc[..., :h], qc[..., h:2 * h], qc[..., 2 * h:], ctx, d_out.cuda(), lse, nh,
                               pivot_idx.cuda(), w, times).float().cpu()
    torch.cuda.synchronize()

    def heads_first(t):
        return t.float().view(b, s, nh, 64).permute(0, 2, 1, 3).contiguous()
    q, k, v = (heads_first(qkv[..., i * h:(i + 1) * h]).requires_grad_(True) for i in range(3))
    pam = _rmask(s, w, times).expand(b, s, s).gather(dim=-1, index=pivot_idx.unsqueeze(1).expand(b, s, n_piv))
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    o = O.sparse_attention(q, k, v, pivot_idx, pam, w, times)
    o.backward(heads_first(d_out))
    o_tok = o.detach().permute(0, 2, 1, 3).reshape(b, s, h)
    err = (ctx.float().cpu() - o_tok).abs().max().item() / o_tok.abs().max().item()
    print("s=%d w=%d x%d piv=%d: fwd rel err %.3e" % (s, w, times, n_piv, err))
    assert err < 2e-2
    for name, got, ref in (("dq", dqkv[..., :h], q.grad), ("dk", dqkv[..., h:2 * h], k.grad),
                           ("dv", dqkv[..., 2 * h:], v.grad)):
        ref_tok = ref.permute(0, 2, 1, 3).reshape(b, s, h)
        e = (got - ref_tok).abs().max().item() / ref_tok.abs().max().item()
        print("   %s rel err %.3e" % (name, e))
        assert e < 4e-2, (name, e)


def test_model_sparse_training_step_matches_oracle():
    """GPT2Model(..., checkpoint_activations=True) forward + loss + backward with is_sparse = 1 against the oracle run
    with the SAME pivots (Python `random` is re-seeded before each side: the model draws fresh pivots per layer in
    the reference's order, mpu/sparse_transformer.py:556-565)."""
    from cogview_b200 import mpu
    from cogview_b200.model import GPT2Model
    cfg = dict(num_layers=2, vocab_size=58240, hidden_size=256, num_attention_heads=4, max_sequence_length=256)
    s, w, times, n_piv = 256, 64, 2, 48
    sd32 = recipes.gpt2_state_dict(seed=21, **cfg)
    m = GPT2Model(num_layers=2, vocab_size=cfg["vocab_size"], hidden_size=256, num_attention_heads=4,
                  embedding_dropout_prob=0.0, attention_dropout_prob=0.0, output_dropout_prob=0.0,
                  max_sequence_length=s, max_memory_length=0, checkpoint_activations=True, checkpoint_num_layers=1,
                  query_window=w, key_window_times=times, num_pivot=n_piv)
    m.load_state_dict(sd32)
    m = m.cuda().bfloat16().train()
    tokens = recipes.text_image_tokens(2, 32, s - 32, seed=3)
    labels = torch.roll(tokens, -1, dims=1)
    pos = torch.arange(s).unsqueeze(0).expand(2, -1).contiguous()
    img = tokens < recipes.IMG_VOCAB
    random.seed(77)
    logits, *_ = m(tokens.cuda(), pos.cuda(), torch.tril(torch.ones((1, 1, s, s), device="cuda")), (~img).cuda(),
                   img.cuda(), 1)
    loss = mpu.vocab_parallel_cross_entropy(logits.contiguous().float(), labels.cuda()).mean()
    loss.backward()
    # oracle with the same per-layer pivots
    random.seed(77)
    img_all = [img[i].nonzero(as_tuple=False).view(-1) for i in range(2)]
    txt_all = [(~img)[i].nonzero(as_tuple=False).view(-1) for i in range(2)]
    from cogview_b200.mpu.sparse_transformer import GPT2ParallelTransformer
    pivots = [GPT2ParallelTransformer.sample_pivot_idx(img_all, txt_all, n_piv) for _ in range(2)]
    sdr = {k: v.to(torch.bfloat16).float().requires_grad_(True) for k, v in sd32.items()}
    rm = _rmask(s, w, times)
    x = torch.nn.functional.embedding(tokens, sdr["word_embeddings.weight"]) + \
        torch.nn.functional.embedding(pos, sdr["transformer.position_embeddings.weight"])
    for li in range(2):
        pam = rm.expand(2, s, s).gather(dim=-1, index=pivots[li].unsqueeze(1).expand(2, s, n_piv))
        x = O.transformer_layer(sdr, li, x, pam, 4, is_sparse=1, pivot_idx=pivots[li], query_window=w,
                                key_window_times=times)
    xf = O.layernorm_absmax(x, sdr["transformer.final_layernorm.weight"], sdr["transformer.final_layernorm.bias"])
    o_logits = torch.nn.functional.linear(xf, sdr["word_embeddings.weight"])
    o_loss = O.vocab_parallel_cross_entropy(o_logits, labels).mean()
    o_loss.backward()
    scale = o_logits.abs().max().item()
    err = (logits.float().cpu() - o_logits.detach()).abs().max().item()
    print("sparse training step: logits max|diff| %.3e (scale %.3e), loss %.5f vs oracle %.5f" % (
        err, scale, loss.item(), o_loss.item()))
    assert err < 2e-2 * scale and abs(loss.item() - o_loss.item()) < 1e-2
    for n, p in m.named_parameters():
        ref = sdr[n].grad
        e = ((p.grad.float().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()
        assert e < 6e-2, (n, e)
