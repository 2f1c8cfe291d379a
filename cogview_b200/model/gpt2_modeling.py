"""GPT-2 model — mirror of the reference's model/gpt2_modeling.py (GPT2Model :55-123,
gpt2_get_params_for_weight_decay_optimization :35-52): same constructor and forward signatures, same
state_dict keys (`word_embeddings.weight`, `transformer.*`)."""
import os

import torch

from .. import mpu, ops
from ..mpu.layers import _as_bf16
from ..mpu.sparse_transformer import _EmbedFn, mask_to_sep


def init_method_normal(std=0.02):
    def init_(tensor):
        return torch.nn.init.normal_(tensor, mean=0.0, std=std)
    return init_


def gpt2_get_params_for_weight_decay_optimization(module):
    """Weights decay; biases and LayerNorm parameters do not (model/gpt2_modeling.py:35-52)."""
    weight_decay_params = {'params': []}
    no_weight_decay_params = {'params': [], 'weight_decay': 0.0}
    for module_ in module.modules():
        if isinstance(module_, (mpu.LayerNorm, torch.nn.LayerNorm)):
            no_weight_decay_params['params'].extend([p for p in list(module_._parameters.values()) if p is not None])
        else:
            weight_decay_params['params'].extend(
                [p for n, p in list(module_._parameters.items()) if p is not None and n != 'bias'])
            no_weight_decay_params['params'].extend(
                [p for n, p in list(module_._parameters.items()) if p is not None and n == 'bias'])
    return weight_decay_params, no_weight_decay_params


class _LogitsFn(torch.autograd.Function):
    """logits = h W^T with the tied word-embedding matrix (model/gpt2_modeling.py:117-118), fp32 output."""

    @staticmethod
    def forward(ctx, hidden, weight):
        w = _as_bf16(weight)
        ctx.save_for_backward(hidden, w)
        ctx.wdtype = weight.dtype
        V = w.shape[0]
        ld = (V + 3) // 4 * 4
        out = torch.empty((hidden.shape[0], ld), dtype=torch.float32, device=hidden.device)
        ops.gemm(hidden, w, out_dtype=torch.float32, out=out[:, :V] if ld != V else out)
        return out[:, :V]

    @staticmethod
    def backward(ctx, d_logits):
        hidden, w = ctx.saved_tensors
        dl = _as_bf16(d_logits)
        if dl.stride(1) != 1 or dl.stride(0) % 8 != 0:
            V = dl.shape[1]
            buf = torch.empty((dl.shape[0], (V + 7) // 8 * 8), dtype=torch.bfloat16, device=dl.device)
            buf[:, :V].copy_(dl)
            dl = buf[:, :V]
        d_hidden = ops.gemm(dl, w, b_mn_major=True)
        d_w = ops.gemm(dl, hidden, a_mn_major=True, b_mn_major=True)
        return d_hidden, d_w.to(ctx.wdtype)


class GPT2Model(torch.nn.Module):
    """GPT-2 language model over concatenated text + image tokens.  forward returns (logits, *mems)."""

    def __init__(self, num_layers, vocab_size, hidden_size, num_attention_heads, embedding_dropout_prob,
                 attention_dropout_prob, output_dropout_prob, max_sequence_length, max_memory_length,
                 checkpoint_activations, checkpoint_num_layers=1, parallel_output=True, query_window=128,
                 key_window_times=6, num_pivot=768):
        super().__init__()
        self.parallel_output = parallel_output
        init_method = init_method_normal(std=0.02)
        self.word_embeddings = mpu.VocabParallelEmbedding(vocab_size, hidden_size, init_method=init_method)
        self.transformer = mpu.GPT2ParallelTransformer(
            num_layers, hidden_size, num_attention_heads, max_sequence_length, max_memory_length,
            embedding_dropout_prob, attention_dropout_prob, output_dropout_prob, checkpoint_activations,
            checkpoint_num_layers, query_window=query_window, key_window_times=key_window_times, num_pivot=num_pivot)

    def forward(self, input_ids, position_ids, attention_mask, txt_indices_bool, img_indices_bool, is_sparse, *mems,
                logits_last_only=False):
        """Same positional arguments as the reference (model/gpt2_modeling.py:106).  Returns
        (logits [b, s, V] fp32, *mems).  `logits_last_only` (keyword, extension): only the last position's
        logits are computed — what the sampling loop reads (generation/sampling.py:155)."""
        tr = self.transformer
        b, sq = input_ids.shape
        mem_len = mems[0].size(1) if mems else 0
        sep = 0 if is_sparse != 0 else mask_to_sep(attention_mask, sq, sq + mem_len)
        if position_ids.shape != input_ids.shape:
            position_ids = position_ids.expand_as(input_ids)
        fast = self._fast_decode(input_ids, position_ids, mems, b, sq) if is_sparse == 0 else None
        if fast is not None:
            return fast
        wte, wpe = self.word_embeddings.weight, tr.position_embeddings.weight
        if torch.is_grad_enabled() and (wte.requires_grad or wpe.requires_grad):
            p_emb = tr.embedding_dropout_prob if tr.training else 0.0
            x, am_x = _EmbedFn.apply(input_ids, position_ids, wte, wpe, p_emb)
        else:
            am_x = ops.new_scalars(1, wte.device)
            x = ops.embed_fwd(input_ids, position_ids, _as_bf16(wte.detach()).contiguous(),
                              _as_bf16(wpe.detach()).contiguous(), am_x)
        y, mem_layers = tr.run_layers(x, am_x, b, sq, sep, mems, is_sparse=is_sparse,
                                      txt_indices_bool=txt_indices_bool, img_indices_bool=img_indices_bool)
        h = y.shape[1]
        if logits_last_only:
            y = y.view(b, sq, h)[:, -1].contiguous()
            sq_out = 1
        else:
            sq_out = sq
        if torch.is_grad_enabled() and (y.requires_grad or wte.requires_grad):
            logits = _LogitsFn.apply(y, wte)
        else:
            logits = _LogitsFn.forward(_NoCtx(), y, wte.detach())
        return (logits.view(b, sq_out, -1), *mem_layers)


def _fast_decode_impl(self, input_ids, position_ids, mems, b, sq):
    """One token per sequence on the K|V cache (the call pattern of generation/sampling.py:147-151): CUDA-graph
    replay of the weight-streaming decode step.  Returns None when the call does not qualify."""
    tr = self.transformer
    if (sq != 1 or not mems or tr.mems_mode != 'kv' or tr.max_memory_length <= 0 or torch.is_grad_enabled()
            or b > 16 or os.environ.get('COGVIEW_B200_FAST_DECODE', '1') == '0'):
        return None
    from ..mpu import kv_cache
    from ..mpu.decode import DecodeRunner
    caches = kv_cache.prepare(tr, mems, b, sq)
    runner = getattr(caches, 'runner', None)
    if runner is None:
        runner = caches.runner = DecodeRunner(self, caches,
                                              use_graph=os.environ.get('COGVIEW_B200_CUDA_GRAPH', '1') != '0')
    logits = runner.step(input_ids, position_ids, caches.t)
    return (logits.view(b, 1, -1), *caches.views())


GPT2Model._fast_decode = _fast_decode_impl


def _generate_run_impl(self, last_tokens, first_pos, mems, n_steps, temperature, top_k, invalid_slices, sparse=None):
    """n_steps sampled tokens in a row on the K|V cache — the inner loop of generation/sampling.py:147-183 for a
    stretch of the template that is all 'generate' slots — with the sampling tail inside the replayed CUDA graph
    (mpu/decode.py sample_run).  Returns None when the fast path does not apply, else
    (tokens [b, n_steps], summed log-probabilities [b], mems).
    sparse = dict(n_img=<image vocabulary size>, tokens=<[b, t + 1] all tokens so far>) runs the stretch with
    is_sparse == 2 semantics (mpu/sparse_transformer.py:498-520, :591-600, :727-750): pivots + trailing window per layer per
    token, chosen on the device (cv_sparse_plan) instead of Python's random.sample — same distribution, other stream."""
    tr = self.transformer
    b = last_tokens.shape[0]
    if (not mems or tr.mems_mode != 'kv' or tr.max_memory_length <= 0 or torch.is_grad_enabled() or b > 16
            or n_steps < 1 or os.environ.get('COGVIEW_B200_FAST_DECODE', '1') == '0'
            or os.environ.get('COGVIEW_B200_GRAPH_SAMPLING', '1') == '0'):
        return None
    from ..mpu import kv_cache
    from ..mpu.decode import DecodeRunner
    if mems[0].size(1) + n_steps > tr.max_memory_length:
        return None
    caches = kv_cache.prepare(tr, mems, b, n_steps)
    runner = getattr(caches, 'runner', None)
    if runner is None:
        runner = caches.runner = DecodeRunner(self, caches,
                                              use_graph=os.environ.get('COGVIEW_B200_CUDA_GRAPH', '1') != '0')
    pos = torch.full((b, 1), int(first_pos), dtype=torch.long, device=last_tokens.device)
    if sparse is not None and (b > 16 or tr.max_memory_length > 4096 or runner.persistent):
        return None
    new_tokens, logp = runner.sample_run(last_tokens.reshape(b, 1), pos, caches.t, n_steps, temperature, top_k,
                                         invalid_slices, sparse=sparse)
    return new_tokens, logp, caches.views()


GPT2Model.generate_run = _generate_run_impl


class _NoCtx:
    def save_for_backward(self, *a):
        pass
