"""Autoregressive token filling — mirror of the reference's generation/sampling.py (`top_k_logits` :24-49,
`get_batch` :51-63, `filling_sequence` :64-186, `shrink_beams` :188-198, `add_interlacing_beam_marks`
:200-211, `inverse_prompt_score` :214-230) and of `get_masks_and_position_ids` (pretrain_gpt2.py:210-253).

Same call signatures and semantics; two host-side differences that do not change results:
  * the template `seq` is read once into a Python list (the reference indexes a CUDA tensor every step, which
    forces a device sync per generated token);
  * the tokenizer is any object with the UnifiedTokenizer surface the loop uses (`tok['[BOI1]']`,
    `tok.img_tokenizer.num_tokens`, `tok.txt_tokenizer.num_tokens`); `TokenLayout` provides that surface from
    the vocabulary sizes alone (the role FakeTokenizer plays in data_utils/unified_tokenizer.py:208-212).
"""
import math
import os as _os

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------
# token layout (data_utils/unified_tokenizer.py:23-67): [image codes | text pieces | 27 command tokens]
# ----------------------------------------------------------------------------------------------------
class _Count:
    def __init__(self, n):
        self.num_tokens = n

    def __len__(self):
        return self.num_tokens


class TokenLayout:
    RAW_COMMAND_TOKENS = ['[PAD]', '[BOI1]', '[BOI2]', '[BOI3]', '[EOI1]', '[EOI2]', '[EOI3]', '[ROI1]', '[ROI2]',
                          '[ROI3]', '[SEP]', '[MASK]', '[CLS]', '[ENC]', '[TINY]', '[SMALL]', '[BASE]', '[BIG]',
                          '[POS0]', '[POS1]', '[POS2]', '[POS3]', '[POS4]', '[POS5]', '[POS6]', '[POS7]', '[POS8]']

    def __init__(self, img_tokens=8192, txt_tokens=50000):
        self.img_tokenizer = _Count(img_tokens)
        self.txt_tokenizer = _Count(txt_tokens)
        base = img_tokens + txt_tokens
        self.command_tokens = {k: base + i for i, k in enumerate(self.RAW_COMMAND_TOKENS)}
        self.num_tokens = base + len(self.RAW_COMMAND_TOKENS)

    def __getitem__(self, command_token):
        return self.command_tokens[command_token]

    def wrap_code(self, code, idx=1):
        """data_utils/unified_tokenizer.py:125-151: [size tag] [BOIidx] code [EOIidx]; the tag follows the side of
        the (square) code grid.  Lists, numpy arrays and tensors keep their type."""
        n = len(code)
        side = int(math.sqrt(n) + 1e-4)
        assert side * side == n, 'image codes must form a square grid'
        prefix = {8: '[TINY]', 16: '[SMALL]', 32: '[BASE]', 64: '[BIG]'}[side]
        head = [self.command_tokens[prefix], self.command_tokens['[BOI%d]' % idx]]
        tail = [self.command_tokens['[EOI%d]' % idx]]
        if isinstance(code, list):
            return head + code + tail
        if isinstance(code, np.ndarray):
            return np.concatenate((np.array(head), code, np.array(tail)), axis=0)
        if isinstance(code, torch.Tensor):
            return torch.cat((torch.tensor(head, dtype=code.dtype, device=code.device), code,
                              torch.tensor(tail, dtype=code.dtype, device=code.device)))
        raise ValueError('code must be a list, numpy array or tensor')

    def __len__(self):
        return self.num_tokens


_TOKENIZER = None


def set_tokenizer(tok):
    global _TOKENIZER
    _TOKENIZER = tok


def get_tokenizer(args=None):
    global _TOKENIZER
    if _TOKENIZER is None:
        n_img = getattr(args, 'img_tokenizer_num_tokens', None) or 8192
        _TOKENIZER = TokenLayout(img_tokens=n_img)
    return _TOKENIZER


# ----------------------------------------------------------------------------------------------------
def get_masks_and_position_ids(data, loss_mask=None, attention_mask=None, args=None):
    """pretrain_gpt2.py:210-253 (pre-training branch): lower-triangular [1,1,s,s] mask, ones loss mask, arange
    position ids."""
    batch_size, seq_length = data.size()
    if attention_mask is None:
        attention_mask = torch.tril(torch.ones((1, seq_length, seq_length), device=data.device)).unsqueeze(1)
    if loss_mask is None:
        loss_mask = torch.ones(data.size(), dtype=torch.float, device=data.device)
    position_ids = torch.arange(seq_length, dtype=torch.long, device=data.device).unsqueeze(0).expand_as(data)
    return attention_mask, loss_mask, position_ids


def top_k_logits(logits, top_k=0, top_p=0.0, filter_value=-float('Inf')):
    """generation/sampling.py:24-49."""
    if top_k > 0:
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits[logits < kth] = filter_value
    if top_p > 0.0:
        logits = logits.view(logits.size()[1]).contiguous()
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative_probs > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        logits[sorted_indices[remove]] = filter_value
        logits = logits.view(1, -1).contiguous()
    return logits


def get_batch(context_tokens, device, args):
    tokens = context_tokens
    tokens = tokens.unsqueeze(0).contiguous() if tokens.dim() == 1 else tokens.view(tokens.shape[0], -1).contiguous()
    tokens = tokens.to(device)
    attention_mask, loss_mask, position_ids = get_masks_and_position_ids(tokens)
    return tokens, attention_mask, position_ids


def shrink_beams(tokens, mems, nb, score):
    """generation/sampling.py:188-198 — keep the best-scoring beam when the beam count shrinks."""
    if tokens.shape[0] == nb:
        return tokens, mems, score
    if torch.is_tensor(score):      # accumulated on the device (no per-token sync); read it only when needed
        score = score.tolist()
    max_idx = score.index(max(score))
    tokens = tokens[max_idx].unsqueeze(0)
    return tokens, [mem[max_idx: max_idx + 1] for mem in mems], [0]


def add_interlacing_beam_marks(seq, nb=12, period=3000):
    """generation/sampling.py:200-211 — replace each -1 (to be generated) by -nb."""
    assert isinstance(seq, list) or len(seq.shape) == 1
    blk_cnt = 0
    for i in range(len(seq)):
        if seq[i] == -1:
            blk_cnt += 1
            seq[i] = -nb
            if blk_cnt == period:
                nb += (nb % 2) * 2 - 1
                blk_cnt = 0
        else:
            blk_cnt = 0


def filling_sequence(model, seq, args, mems=None, invalid_slices=[], **kwargs):
    """generation/sampling.py:64-186.  seq: 1-D LongTensor template [ctx..., -N (generate with N beams), ...].
    Returns the filled token rows [nb, len(seq)]."""
    tokenizer = get_tokenizer(args)
    device = seq.device
    assert len(seq.shape) == 1
    tmpl = seq.tolist()
    out_seq_length = len(tmpl)
    n_img = tokenizer.img_tokenizer.num_tokens
    n_txt = tokenizer.txt_tokenizer.num_tokens
    boi = (tokenizer['[BOI1]'], tokenizer['[BOI2]'])
    eoi = (tokenizer['[EOI1]'], tokenizer['[EOI2]'])
    roi2 = tokenizer['[ROI2]']
    offset = 100000
    invalid_slices = [slice(0, n_img)]

    def slices_after(tok, cur):
        if tok in boi:
            return [slice(n_img, None)]
        if tok in eoi:
            return [slice(0, n_img), slice(n_img + n_txt, None)]
        return cur

    context_length = 0
    while context_length < out_seq_length and tmpl[context_length] >= 0:
        invalid_slices = slices_after(tmpl[context_length], invalid_slices)
        if tmpl[context_length] == roi2:
            offset = context_length
        context_length += 1
    tokens, attention_mask, position_ids = get_batch(seq[:context_length], device, args)

    counter = context_length - 1
    index = 0
    if mems is None:
        mems = []
    score = [0]
    is_sparse = getattr(args, 'is_sparse', 0)
    if is_sparse == 2:
        img_indices_bool = tokens < n_img
        txt_indices_bool = ~img_indices_bool
    elif is_sparse == 0:
        txt_indices_bool = img_indices_bool = None
    else:
        raise ValueError('set is_sparse==2 for inference.')

    while counter < out_seq_length - 1:
        nxt = tmpl[counter + 1]
        invalid_slices = slices_after(nxt, invalid_slices)
        if index == 0:      # first call: the whole context
            position_ids = position_ids.clone()
            position_ids[position_ids > offset] -= offset
            logits, *mems = model(tokens, position_ids, attention_mask, txt_indices_bool, img_indices_bool, is_sparse,
                                  *mems)
            index = counter
        elif nxt >= 0:      # provided token
            if nxt == roi2:
                offset = counter + 1
            tokens, mems, score = shrink_beams(tokens, mems, 1, score)
            counter += 1
            tokens = torch.cat((tokens, seq[counter: counter + 1].expand(tokens.shape[0], 1)), dim=1)
            if is_sparse == 2:
                img_indices_bool = tokens < n_img
                txt_indices_bool = ~img_indices_bool
            continue
        else:
            assert tokens.shape[1] == counter + 1
            # A stretch of identical 'generate' slots: run it with the sampling tail inside the decode graph (no host
            # work per token).  Same operations as the loop body below; falls through when it does not apply.
            run = 1
            while counter + 1 + run < out_seq_length and tmpl[counter + 1 + run] == nxt:
                run += 1
            first_pos = counter - offset if counter > offset else counter
            device_pivots = is_sparse == 2 and _os.environ.get('COGVIEW_B200_SPARSE_PIVOTS', 'device') != 'host'
            if (run >= 2 and index == counter and tokens.shape[0] == -nxt and args.top_p == 0.0
                    and (is_sparse == 0 or device_pivots)
                    and not (counter <= offset < counter + run) and hasattr(model, 'generate_run')):
                sparse = dict(n_img=n_img, tokens=tokens) if is_sparse == 2 else None
                res = model.generate_run(tokens[:, counter:], first_pos, mems, run, args.temperature, args.top_k,
                                         invalid_slices, sparse=sparse)
                if res is not None:
                    new_tokens, logp, mems = res
                    if -nxt > 1:
                        score = (score if torch.is_tensor(score) else torch.tensor(score, device=device)) + logp
                    tokens = torch.cat((tokens, new_tokens), dim=1)
                    if is_sparse == 2:
                        img_indices_bool = tokens < n_img
                        txt_indices_bool = ~img_indices_bool
                    counter += run
                    index = counter
                    continue
            position_ids = torch.arange(index, counter + 1, dtype=torch.long, device=device).unsqueeze(0)
            position_ids[position_ids > offset] -= offset
            tokens, mems, score = shrink_beams(tokens, mems, -nxt, score)
            logits, *mems = model(tokens[:, index:], position_ids, 0, txt_indices_bool, img_indices_bool, is_sparse,
                                  *mems)
            index = counter
        nb = -nxt
        counter += 1
        index += 1

        logits = logits[:, -1]
        logits /= args.temperature
        for sl in invalid_slices:
            logits[..., sl] = -float('Inf')
        logits = top_k_logits(logits, top_k=args.top_k, top_p=args.top_p)
        log_probs = F.softmax(logits, dim=-1)

        if nb > 1 and tokens.shape[0] == 1:     # 1 -> nb beams
            tokens = tokens.expand(nb, -1).contiguous()
            mems = [mem.expand(nb, -1, -1) for mem in mems]
            prev = torch.multinomial(log_probs, num_samples=nb, replacement=True)
            score = torch.log(torch.gather(log_probs, dim=1, index=prev)[0])
        else:
            assert tokens.shape[0] == nb
            prev = torch.multinomial(log_probs, num_samples=1)
            if nb > 1:      # beam scores are only consulted when beams shrink: keep them on the device
                score_plus = torch.log(torch.gather(log_probs, dim=1, index=prev)[:, 0])
                score = (score if torch.is_tensor(score) else torch.tensor(score, device=device)) + score_plus
        tokens = torch.cat((tokens, prev.view(tokens.shape[0], 1)), dim=1)
        if is_sparse == 2:
            img_indices_bool = tokens < n_img
            txt_indices_bool = ~img_indices_bool
    return tokens.view(tokens.shape[0], -1).contiguous()


def inverse_prompt_score(model, seq, args):
    """generation/sampling.py:214-230 — caption log-likelihood given the image (post-selection)."""
    tokenizer = get_tokenizer(args)
    device = seq.device
    assert len(seq.shape) == 2
    botext = 2 + 1024 + 1
    assert tokenizer['[ROI1]'] == seq[0][botext]
    tokens, attention_mask, position_ids = get_batch(seq, device, args)
    logits, *mems = model(tokens, position_ids, attention_mask, None, None, getattr(args, 'is_sparse', 0))
    logits[..., :tokenizer.img_tokenizer.num_tokens] = -float('Inf')
    log_probs = torch.log(F.softmax(logits, dim=-1))
    pred = log_probs[:, botext:-1, :]
    target = tokens[:, botext + 1:].unsqueeze(-1)
    return torch.gather(pred, dim=2, index=target).squeeze(-1).sum(dim=-1)
