"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's hot-path arithmetic.

This file is the oracle: a plain, functional PyTorch-fp32 restatement of what THUDM/CogView computes on
the path this repo accelerates.  It is NOT product code: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  It travels to the GPU box (the
reference tree does not).  Pinning: oracle/make_golden.py runs the UNMODIFIED reference modules
(oracle/ref_harness.py) in the build container, asserts this restatement reproduces them (fp32,
<= 2e-5 abs) and writes tests/golden/*.npz; the reference itself ships no golden vectors or tests for
this path, so those fixtures are the pin.

Parameters are passed as a state_dict with the reference's key names:
  word_embeddings.weight, transformer.position_embeddings.weight,
  transformer.layers.{i}.{input_layernorm,post_attention_layernorm,third_layernorm,fourth_layernorm}.{weight,bias},
  transformer.layers.{i}.attention.{query_key_value,dense}.{weight,bias},
  transformer.layers.{i}.mlp.{dense_h_to_4h,dense_4h_to_h}.{weight,bias}, transformer.final_layernorm.{weight,bias}
Autograd through these functions is the backward oracle.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1.0e-5


# ------------------------------------------------------------------------------------------------
# transformer pieces
# ------------------------------------------------------------------------------------------------
def layernorm_absmax(x, weight, bias, eps=LN_EPS):
    """mpu/sparse_transformer.py:40-44 — LayerNorm of x / (max|x| / 8); the max is over the WHOLE tensor
    and is detached."""
    c = x.detach().abs().max() / 8
    return F.layer_norm(x / c, (x.shape[-1],), weight, bias, eps)


def gelu(x):
    """mpu/sparse_transformer.py:172-176 (OpenAI tanh GELU)."""
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def build_sep_mask(query_length, key_length, sep, dtype=torch.float32):
    """mpu/sparse_transformer.py:482-488 — int `sep`: keys [0, sep + mem) fully visible, rest causal."""
    m = torch.ones((1, query_length, key_length), dtype=dtype)
    m[0, :, -query_length:] = torch.tril(m[0, :, -query_length:])
    m[0, :, :sep + (key_length - query_length)] = 1
    return m.unsqueeze(1)


def standard_attention(q, k, v, mask):
    """mpu/sparse_transformer.py:652-673 — q pre-scaled by 1/sqrt(hn); scores*mask - 10000*(1-mask);
    softmax; no dropout (parity runs use p = 0).  q,k,v: [b, np, s, hn]; mask broadcastable [1,1,sq,sk]."""
    if mask.dim() == 3:
        mask = mask.unsqueeze(1)
    scores = torch.matmul(q / math.sqrt(q.shape[-1]), k.transpose(-1, -2))
    scores = scores * mask - 10000.0 * (1.0 - mask)
    probs = torch.softmax(scores, dim=-1)
    return torch.matmul(probs, v)


def _overlapping_windows(x, w, times):
    """mpu/sparse_transformer.py:629-650 (_chunk): left-pad by (times-1)*w, then windows of times*w keys
    starting every w.  x: [b, np, s, hn] -> [b, np, s//w, times*w, hn]."""
    b, n, s, hn = x.shape
    assert s % w == 0
    xp = F.pad(x, (0, 0, (times - 1) * w, 0))
    g = s // w
    idx = (torch.arange(g).unsqueeze(1) * w + torch.arange(times * w).unsqueeze(0)).reshape(-1)
    return xp[:, :, idx].reshape(b, n, g, times * w, hn)


def sparse_attention(q, k, v, pivot_idx, pivot_attention_mask, query_window=128, key_window_times=6):
    """mpu/sparse_transformer.py:675-725 — training-mode sparse attention: joint softmax over gathered pivot
    keys (bias +log(s // n_piv), masked by the gathered rmask) and a causal band of `key_window_times`
    windows.  pivot_idx: [b, n_piv] int64; pivot_attention_mask: [b, s, n_piv]."""
    b, n_head, s, hn = q.shape
    n_piv = pivot_idx.shape[1]
    w, times = query_window, key_window_times
    gidx = pivot_idx.view(b, 1, n_piv, 1).expand(b, n_head, n_piv, hn)
    pk, pv = torch.gather(k, 2, gidx), torch.gather(v, 2, gidx)
    pm = pivot_attention_mask.unsqueeze(1)
    sp = torch.matmul(q, pk.transpose(-1, -2)) * (pm / math.sqrt(hn)) - 10000.0 * (1.0 - pm)
    sp = sp + math.log(s // n_piv)
    if s % w != 0:
        raise ValueError('The seq_len must be exactly divided by window_size.')
    wk, wv = _overlapping_windows(k, w, times), _overlapping_windows(v, w, times)
    wq = q.view(b, n_head, s // w, w, hn)
    sw = torch.matmul(wq, wk.transpose(-1, -2))
    wm = torch.ones((w, w * times), dtype=sw.dtype).tril_(diagonal=w * (times - 1))
    sw = sw * (wm / math.sqrt(hn)) - 10000.0 * (1.0 - wm)
    sw = sw.clone()
    for t in range(1, times):  # left padding of the first windows
        sw[:, :, t - 1, :, :w * times - w * t] -= 10000.0
    sw = sw.view(b, n_head, s, w * times)
    probs = torch.softmax(torch.cat((sp, sw), dim=-1), dim=-1)
    ctx_p = torch.matmul(probs[..., :n_piv], pv)
    ctx_w = torch.einsum('bcgwk,bcgkh->bcgwh', probs[..., n_piv:].view(b, n_head, s // w, w, w * times), wv)
    return ctx_p + ctx_w.reshape(b, n_head, s, hn)


def sparse_attention_inference(q, k, v, pivot_and_window_idx):
    """mpu/sparse_transformer.py:727-750 — dense softmax over K[idx], V[idx]; the last sq indices are the
    queries themselves (causal among them)."""
    b, n_head, sq, hn = q.shape
    n = pivot_and_window_idx.shape[1]
    gidx = pivot_and_window_idx.view(b, 1, n, 1).expand(b, n_head, n, hn)
    pk, pv = torch.gather(k, 2, gidx), torch.gather(v, 2, gidx)
    scores = torch.matmul(q / math.sqrt(hn), pk.transpose(-1, -2))
    if sq > 1:
        m = torch.full((sq, sq), -10000.0, dtype=q.dtype).triu_(diagonal=1)
        scores = scores.clone()
        scores[:, :, -sq:, -sq:] += m
    return torch.matmul(torch.softmax(scores, dim=-1), pv)


def _heads(t, n_heads):
    b, s, h = t.shape
    return t.view(b, s, n_heads, h // n_heads).permute(0, 2, 1, 3)


def self_attention(sd, pre, x, mask, n_heads, mem=None, is_sparse=0, pivot_idx=None, query_window=128,
                   key_window_times=6):
    """mpu/sparse_transformer.py:123-169.  x: LN'd hidden [b, s, h]; mem: LN'd memory [b, t, h] or None."""
    sq = x.shape[1]
    src = x if mem is None else torch.cat((mem, x), 1)
    mixed = F.linear(src, sd[pre + 'query_key_value.weight'], sd[pre + 'query_key_value.bias'])
    qm, km, vm = mixed.chunk(3, dim=-1)
    qm = qm[:, -sq:]
    q, k, v = _heads(qm, n_heads), _heads(km, n_heads), _heads(vm, n_heads)
    if is_sparse == 1:
        ctx = sparse_attention(q, k, v, pivot_idx, mask, query_window, key_window_times)
    elif is_sparse == 2:
        ctx = sparse_attention_inference(q, k, v, pivot_idx)
    else:
        ctx = standard_attention(q, k, v, mask)
    b, _, _, hn = ctx.shape
    ctx = ctx.permute(0, 2, 1, 3).reshape(b, sq, n_heads * hn)
    return F.linear(ctx, sd[pre + 'dense.weight']) + sd[pre + 'dense.bias']


def mlp(sd, pre, x):
    """mpu/sparse_transformer.py:226-234."""
    h = gelu(F.linear(x, sd[pre + 'dense_h_to_4h.weight'], sd[pre + 'dense_h_to_4h.bias']))
    return F.linear(h, sd[pre + 'dense_4h_to_h.weight']) + sd[pre + 'dense_4h_to_h.bias']


def transformer_layer(sd, i, x, mask, n_heads, mem=None, **attn_kw):
    """mpu/sparse_transformer.py:314-342 — Sandwich-LN block."""
    p = 'transformer.layers.%d.' % i

    def ln(name, t):
        return layernorm_absmax(t, sd[p + name + '.weight'], sd[p + name + '.bias'])

    ln1 = ln('input_layernorm', x)
    mem_n = ln('input_layernorm', mem) if mem is not None else None
    att = self_attention(sd, p + 'attention.', ln1, mask, n_heads, mem=mem_n, **attn_kw)
    att = ln('third_layernorm', att)
    y = x + att
    m = mlp(sd, p + 'mlp.', ln('post_attention_layernorm', y))
    m = ln('fourth_layernorm', m)
    return y + m


def num_layers_of(sd):
    n = 0
    while ('transformer.layers.%d.input_layernorm.weight' % n) in sd:
        n += 1
    return n


def gpt2_forward(sd, n_heads, input_ids, position_ids, attention_mask, mems=(), max_memory_length=0,
                 return_hiddens=False):
    """model/gpt2_modeling.py:106-123 + mpu/sparse_transformer.py:471-626 (dense path, is_sparse=0).
    attention_mask: [1,1,sq,sk] tensor or int `sep`.  Returns (logits, new_mems) (+ per-layer inputs)."""
    L = num_layers_of(sd)
    x = F.embedding(input_ids, sd['word_embeddings.weight'])
    sq = x.shape[1]
    mem_len = mems[0].shape[1] if len(mems) else 0
    if isinstance(attention_mask, int) or attention_mask.numel() == 1:
        attention_mask = build_sep_mask(sq, sq + mem_len, int(attention_mask), x.dtype)
    x = x + F.embedding(position_ids, sd['transformer.position_embeddings.weight'])
    hiddens = [x.detach()]
    for i in range(L):
        x = transformer_layer(sd, i, x, attention_mask, n_heads, mem=mems[i] if len(mems) else None)
        hiddens.append(x.detach())
    out = layernorm_absmax(x, sd['transformer.final_layernorm.weight'], sd['transformer.final_layernorm.bias'])
    logits = F.linear(out, sd['word_embeddings.weight'])
    new_mems = []
    if max_memory_length > 0:  # update_mems, mpu/sparse_transformer.py:615-626
        new_len = min(max_memory_length, mem_len + sq)
        for i, h in enumerate(hiddens):
            if new_len <= sq:
                new_mems.append(h[:, -new_len:])
            else:
                new_mems.append(torch.cat((mems[i][:, -new_len + sq:], h), dim=1))
    if return_hiddens:
        return logits, new_mems, hiddens
    return logits, new_mems


def vocab_parallel_cross_entropy(logits, target):
    """mpu/cross_entropy.py:27-81 at model-parallel size 1: log(sum exp(l - max)) - (l[target] - max)."""
    logits = logits.float()
    mx = logits.max(dim=-1)[0]
    shifted = logits - mx.unsqueeze(-1)
    sum_exp = shifted.exp().sum(dim=-1)
    pred = torch.gather(shifted, -1, target.unsqueeze(-1)).squeeze(-1)
    return torch.log(sum_exp) - pred


def weighted_loss(losses, tokens, loss_mask, img_vocab=8192, txt_loss_scale=1.0):
    """pretrain_gpt2.py:305-321 — text positions weighted by txt_loss_scale; sum(loss*mask)/sum(mask)."""
    img = tokens < img_vocab
    txt = (~img) & (loss_mask > 0)
    lm = loss_mask.clone().float()
    lm[txt] *= txt_loss_scale
    lm = lm.view(-1)
    return torch.sum(losses.view(-1) * lm) / lm.sum()


def top_k_logits(logits, top_k=0, filter_value=-float('inf')):
    """generation/sampling.py:24-31 (top-k branch)."""
    if top_k > 0:
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, filter_value)
    return logits


# ------------------------------------------------------------------------------------------------
# VQ-VAE (vqvae.api.new_model(): channel=512, embed_dim=256, n_embed=8192, stride=6, n_res_block=0)
# keys: enc_b.blocks.{0,2,4,6}.{weight,bias}, quantize_t.embed, dec.blocks.{0,2,4,6}.{weight,bias}
# ------------------------------------------------------------------------------------------------
def vq_encoder(sd, img):
    """vqvae/vqvae_zc.py:117-164 (stride 6, simple) — 3 x conv k4 s2 p1 with ReLU, ReLU, conv 1x1; NHWC out."""
    x = F.conv2d(img, sd['enc_b.blocks.0.weight'], sd['enc_b.blocks.0.bias'], stride=2, padding=1).relu()
    x = F.conv2d(x, sd['enc_b.blocks.2.weight'], sd['enc_b.blocks.2.bias'], stride=2, padding=1).relu()
    x = F.conv2d(x, sd['enc_b.blocks.4.weight'], sd['enc_b.blocks.4.bias'], stride=2, padding=1).relu()
    x = F.conv2d(x, sd['enc_b.blocks.6.weight'], sd['enc_b.blocks.6.bias'])
    return x.permute(0, 2, 3, 1)


def vq_distances(flat, embed):
    """vqvae/vqvae_zc.py:43-47 — ||z||^2 - 2 z.E + ||E||^2, E: [dim, n_embed]."""
    return flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embed + embed.pow(2).sum(0, keepdim=True)


def vq_quantize(sd, z):
    """vqvae/vqvae_zc.py:41-54,84-86,93 — nearest code (argmax of -dist, first index on ties), lookup."""
    embed = sd['quantize_t.embed']
    flat = z.reshape(-1, embed.shape[0])
    _, ind = (-vq_distances(flat, embed)).max(1)
    ind = ind.view(*z.shape[:-1])
    quant = F.embedding(ind, embed.t())
    diff = (quant - z).pow(2).mean()
    return quant, diff, ind


def vq_decoder(sd, quant_nchw):
    """vqvae/vqvae_zc.py:167-214 (stride 4, simple) — convT k4 s2 p1 x3 with ReLU, conv 1x1 512->3."""
    x = F.conv_transpose2d(quant_nchw, sd['dec.blocks.0.weight'], sd['dec.blocks.0.bias'], stride=2, padding=1).relu()
    x = F.conv_transpose2d(x, sd['dec.blocks.2.weight'], sd['dec.blocks.2.bias'], stride=2, padding=1).relu()
    x = F.conv_transpose2d(x, sd['dec.blocks.4.weight'], sd['dec.blocks.4.bias'], stride=2, padding=1).relu()
    return F.conv2d(x, sd['dec.blocks.6.weight'], sd['dec.blocks.6.bias'])


IMG_STD = (0.30379, 0.32279, 0.32800)
IMG_MEAN = (0.79093, 0.76271, 0.75340)


def img2code(sd, img):
    """vqvae/api.py:22-30."""
    with torch.no_grad():
        _, _, ind = vq_quantize(sd, vq_encoder(sd, img))
    return ind.view(img.shape[0], -1)


def code2img(sd, code):
    """vqvae/api.py:32-44 — code [b,h,w] (or [b,h*w] for b == 1); de-normalised output."""
    if code.dim() == 2:
        s = int(math.sqrt(code.numel()) + 1e-5)
        code = code.view(code.shape[0], s, s)
    with torch.no_grad():
        quant = F.embedding(code, sd['quantize_t.embed'].t()).permute(0, 3, 1, 2)
        out = vq_decoder(sd, quant)
        out = out * torch.tensor(IMG_STD).view(1, -1, 1, 1) + torch.tensor(IMG_MEAN).view(1, -1, 1, 1)
    return out
