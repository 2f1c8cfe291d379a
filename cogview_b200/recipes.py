"""Deterministic synthetic weights and inputs shared by bench.py, smoke(), the golden-vector generator and the
parity tests (a neutral module: the product arm of bench.py must not import oracle/).  Weights are drawn from a
seeded CPU torch.Generator in a fixed key order (identical on every machine with this torch build), so fixtures only
need to store OUTPUTS.

Distributions follow the reference initialisers (normal(0, 0.02); out-proj / 4h->h scaled by
1/sqrt(2*num_layers): mpu/sparse_transformer.py:344-358, 420-423; model/gpt2_modeling.py:23-32), except
that biases and LayerNorm affine parameters are perturbed away from 0/1 when `perturb` is set so that a
parity test exercises them.
"""
import math

import torch

CONFIG1 = dict(num_layers=2, vocab_size=58240, hidden_size=256, num_attention_heads=4, max_sequence_length=128)
COGVIEW_4B = dict(num_layers=48, vocab_size=58240, hidden_size=2560, num_attention_heads=40, max_sequence_length=1089)
IMG_VOCAB = 8192


def gpt2_state_dict(num_layers, vocab_size, hidden_size, max_sequence_length, seed=1234, perturb=True,
                    dtype=torch.float32, **_):
    g = torch.Generator().manual_seed(seed)
    h = hidden_size

    def normal(shape, std):
        return (torch.randn(shape, generator=g) * std).to(dtype)

    def bias(n):
        return normal((n,), 0.02) if perturb else torch.zeros(n, dtype=dtype)

    def ln(prefix, sd):
        sd[prefix + '.weight'] = (1.0 + (normal((h,), 0.1) if perturb else torch.zeros(h))).to(dtype)
        sd[prefix + '.bias'] = normal((h,), 0.1) if perturb else torch.zeros(h, dtype=dtype)

    sd = {}
    sd['word_embeddings.weight'] = normal((vocab_size, h), 0.02)
    sd['transformer.position_embeddings.weight'] = normal((max_sequence_length, h), 0.02)
    out_std = 0.02 / math.sqrt(2.0 * num_layers)
    for i in range(num_layers):
        p = 'transformer.layers.%d.' % i
        ln(p + 'input_layernorm', sd)
        sd[p + 'attention.query_key_value.weight'] = normal((3 * h, h), 0.02)
        sd[p + 'attention.query_key_value.bias'] = bias(3 * h)
        sd[p + 'attention.dense.weight'] = normal((h, h), out_std)
        sd[p + 'attention.dense.bias'] = bias(h)
        ln(p + 'post_attention_layernorm', sd)
        ln(p + 'third_layernorm', sd)
        ln(p + 'fourth_layernorm', sd)
        sd[p + 'mlp.dense_h_to_4h.weight'] = normal((4 * h, h), 0.02)
        sd[p + 'mlp.dense_h_to_4h.bias'] = bias(4 * h)
        sd[p + 'mlp.dense_4h_to_h.weight'] = normal((h, 4 * h), out_std)
        sd[p + 'mlp.dense_4h_to_h.bias'] = bias(h)
    ln('transformer.final_layernorm', sd)
    return sd


def text_image_tokens(batch, n_text, n_image, seed=0, vocab_size=58240):
    """text ids uniform in [8192, 58192), image ids uniform in [0, 8192)."""
    g = torch.Generator().manual_seed(seed)
    txt = torch.randint(IMG_VOCAB, min(58192, vocab_size), (batch, n_text), generator=g)
    img = torch.randint(0, IMG_VOCAB, (batch, n_image), generator=g)
    return torch.cat((txt, img), dim=1)


def vqvae_state_dict(seed=0, channel=512, embed_dim=256, n_embed=8192, dtype=torch.float32):
    """Shapes of vqvae.api.new_model() (vqvae/api.py:12-20); uniform(+-1/sqrt(fan_in)) like torch's conv default."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return ((torch.rand(shape, generator=g) * 2 - 1) * b).to(dtype)

    sd = {}
    sd['enc_b.blocks.0.weight'] = uni((channel, 3, 4, 4), 3 * 16)
    sd['enc_b.blocks.0.bias'] = uni((channel,), 3 * 16)
    sd['enc_b.blocks.2.weight'] = uni((channel, channel, 4, 4), channel * 16)
    sd['enc_b.blocks.2.bias'] = uni((channel,), channel * 16)
    sd['enc_b.blocks.4.weight'] = uni((channel, channel, 4, 4), channel * 16)
    sd['enc_b.blocks.4.bias'] = uni((channel,), channel * 16)
    sd['enc_b.blocks.6.weight'] = uni((embed_dim, channel, 1, 1), channel)
    sd['enc_b.blocks.6.bias'] = uni((embed_dim,), channel)
    bound = (5.0 / 3.0) * math.sqrt(6.0 / (embed_dim + n_embed))  # xavier_uniform(gain=tanh), vqvae_zc.py:35-36
    # codebook scaled up so that nearest-code gaps are well above fp32 noise for the synthetic encoder output
    sd['quantize_t.embed'] = ((torch.rand((embed_dim, n_embed), generator=g) * 2 - 1) * bound).to(dtype)
    sd['quantize_t.cluster_size'] = torch.zeros(n_embed, dtype=dtype)
    sd['quantize_t.embed_avg'] = sd['quantize_t.embed'].clone()
    # ConvTranspose2d weight layout: [Cin, Cout, kh, kw]
    sd['dec.blocks.0.weight'] = uni((embed_dim, channel, 4, 4), channel * 16)
    sd['dec.blocks.0.bias'] = uni((channel,), channel * 16)
    sd['dec.blocks.2.weight'] = uni((channel, channel, 4, 4), channel * 16)
    sd['dec.blocks.2.bias'] = uni((channel,), channel * 16)
    sd['dec.blocks.4.weight'] = uni((channel, channel, 4, 4), channel * 16)
    sd['dec.blocks.4.bias'] = uni((channel,), channel * 16)
    sd['dec.blocks.6.weight'] = uni((3, channel, 1, 1), channel)
    sd['dec.blocks.6.bias'] = uni((3,), channel)
    return sd


def images(batch, size=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((batch, 3, size, size), generator=g)
