"""Callers and data formats either side of the hot path, host logic only: the compact binary
dataset and text+code template (data_utils/datasets.py:63-128, data_utils/templates.py:52-65,
data_utils/unified_tokenizer.py:125-151), batch assembly (pretrain_gpt2.py:256-289), loss weighting (:302-329,
pinned to the reference through the oracle's golden loss) and the magnify window driver (generation/magnify.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import cogview_oracle as O
from oracle import recipes


def _tok():
    from cogview_b200.generation import sampling
    return sampling.TokenLayout()


def test_wrap_code_and_text_code_template():
    from cogview_b200 import data
    tok = _tok()
    code = np.arange(1024) % 8192
    text = np.array([9000, 9001, 9002])
    row = data.TextCodeTemplate(text, code, tok)
    assert row.tolist() == [tok['[ROI1]'], 9000, 9001, 9002, tok['[BASE]'], tok['[BOI1]']] + code.tolist() + [tok['[EOI1]']]
    assert tok.wrap_code(list(range(64)), idx=2)[:2] == [tok['[TINY]'], tok['[BOI2]']]
    t = tok.wrap_code(torch.arange(256), idx=3)
    assert t[0] == tok['[SMALL]'] and t[1] == tok['[BOI3]'] and t[-1] == tok['[EOI3]'] and t.shape == (259,)
    with pytest.raises(AssertionError):
        tok.wrap_code(list(range(100 + 1)))                      # not a square grid


def test_compact_binary_dataset_rows_and_batch(tmp_path):
    from cogview_b200 import data
    tok = _tok()
    g = np.random.default_rng(0)
    rows = np.full((5, 64 + 1024), -1, dtype=np.int32)
    n_text = [3, 64, 10, 1, 30]
    for r, n in enumerate(n_text):
        rows[r, :n] = g.integers(8192, 58192, n)
        rows[r, 64:] = g.integers(0, 8192, 1024)
    path = tmp_path / "merge.bin"
    rows.tofile(str(path))
    max_len = 1089
    for preload in (False, True):
        ds = data.BinaryDataset(str(path), data.compact_binary_process_fn(max_len, tok), preload=preload)
        assert len(ds) == 5
        for r, n in enumerate(n_text):
            item = ds[r]
            valid = 1 + n + 2 + 1024 + 1                          # [ROI1] text [BASE] [BOI1] codes [EOI1]
            assert item['text'].shape == (max_len,) and item['loss_mask'].shape == (max_len,)
            assert item['text'][0] == tok['[ROI1]'] and item['text'][1 + n] == tok['[BASE]']
            assert np.array_equal(item['text'][1:1 + n], rows[r, :n])
            fit = min(1024, max_len - (3 + n))                    # 64 text tokens overflow 1089: the row is truncated
            assert np.array_equal(item['text'][3 + n:3 + n + fit], rows[r, 64:64 + fit])
            if valid <= max_len:
                assert item['text'][valid - 1] == tok['[EOI1]'] and (item['text'][valid:] == tok['[PAD]']).all()
            assert item['loss_mask'].sum() == min(valid, max_len) and item['loss_mask'][:min(valid, max_len)].all()
    ds = data.BinaryDataset(str(path), data.compact_binary_process_fn(max_len, tok))
    batch = torch.utils.data.default_collate([ds[0], ds[2]])
    tokens, labels, loss_mask, attention_mask, position_ids = data.make_batch(batch)
    assert tokens.shape == labels.shape == loss_mask.shape == position_ids.shape == (2, max_len - 1)
    assert torch.equal(tokens[:, 1:], labels[:, :-1]) and tokens.dtype == torch.int64
    assert attention_mask.shape == (1, 1, max_len - 1, max_len - 1)
    assert torch.equal(position_ids[0], torch.arange(max_len - 1))
    assert loss_mask[0].sum().item() == (1 + 3 + 2 + 1024 + 1) - 1               # shifted by one with the labels
    short = data.tokenized_process_fn(16, tok)(np.arange(40).reshape(4, 10))
    assert short['text'].tolist() == list(range(16)) and short['loss_mask'].sum() == 16


def test_weighted_loss_matches_reference_pinned_oracle(golden_dir):
    from cogview_b200 import pretrain
    g = np.load(os.path.join(str(golden_dir), "gpt2_config1.npz"))
    losses = torch.from_numpy(g["losses"])
    tokens = torch.from_numpy(g["tokens_full"])[:, :-1]
    scale = float(g["txt_loss_scale"])
    ones = torch.ones_like(tokens, dtype=torch.float)
    loss, img_loss, txt_loss = pretrain.weighted_loss(losses, tokens, ones, scale, recipes.IMG_VOCAB)
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))    # the reference's number
    assert torch.allclose(loss, O.weighted_loss(losses, tokens, ones, recipes.IMG_VOCAB, scale), atol=1e-6)
    img = tokens < recipes.IMG_VOCAB
    assert torch.allclose(img_loss, losses[img].mean(), atol=1e-5)
    assert torch.allclose(txt_loss, losses[~img].mean(), atol=1e-5)                    # reported without the scale
    # padded positions (mask 0) are neither text nor counted
    mask = ones.clone()
    mask[:, -20:] = 0
    loss2, _, txt2 = pretrain.weighted_loss(losses, tokens, mask, scale, recipes.IMG_VOCAB)
    assert torch.allclose(loss2, O.weighted_loss(losses, tokens, mask, recipes.IMG_VOCAB, scale), atol=1e-6)
    assert torch.allclose(txt2, losses[(~img) & (mask > 0)].mean(), atol=1e-5)


def test_magnify_window_assembly():
    """Every one of the 64 x 64 target codes is generated exactly once; later windows receive the rows produced by
    earlier ones as context; the context is text + 16 x 16 source patch + the five midfix tokens."""
    from cogview_b200.generation import magnify as mg
    tok = _tok()
    code = torch.arange(1024)
    text = torch.tensor([tok['[ROI1]'], 9000, 9001])
    calls, generated = [], torch.zeros(64 * 64, dtype=torch.long)

    def fake_fill(model, seq, args, invalid_slices=None):
        assert invalid_slices == [slice(tok.img_tokenizer.num_tokens, None)]
        ctx = len(text) + 256 + 5
        i, j, line = mg.WINDOWS[len(calls)]
        assert seq.shape[0] == ctx + line * 32
        assert torch.equal(seq[:len(text)], text)
        assert torch.equal(seq[len(text):len(text) + 256], code.view(32, 32)[8 * i:8 * i + 16, 8 * j:8 * j + 16].reshape(-1))
        assert seq[ctx - 5:ctx].tolist() == [tok['[EOI1]'], tok['[ROI2]'], tok['[POS0]'], tok['[BASE]'], tok['[BOI2]']]
        part = seq[ctx:].clone().view(line, 32)
        rows = torch.arange(16 * i, 16 * i + line).view(-1, 1)
        cols = torch.arange(16 * j, 16 * j + 32).view(1, -1)
        ident = 10000 + rows * 64 + cols                                  # what the model "generates" at (row, col)
        todo = part < 0
        generated.index_add_(0, (rows * 64 + cols)[todo], torch.ones(int(todo.sum()), dtype=torch.long))
        assert torch.equal(part[~todo], ident[~todo])                    # context rows come from earlier windows
        calls.append((i, j, line, int(todo.sum())))
        return torch.cat((seq[:ctx], torch.where(todo, ident, part).reshape(-1))).unsqueeze(0)

    out = mg.magnify(None, tok, code, text, None, fill=fake_fill)
    assert out.shape == (1, 4096) and len(calls) == 9
    assert torch.equal(out.view(64, 64), 10000 + torch.arange(4096).view(64, 64))
    assert (generated == 1).all()


def test_train_step_glue_with_a_stand_in_model(monkeypatch):
    """pretrain.train_step control flow (pretrain_gpt2.py:406-450) with a CPU stand-in for the model and the oracle's
    cross-entropy in place of the CUDA op: parameters move, the scheduler steps, and a non-finite forward skips the
    backward / optimizer step and reports skipped_iter = 1."""
    from cogview_b200 import pretrain
    monkeypatch.setattr(pretrain.mpu, "vocab_parallel_cross_entropy", O.vocab_parallel_cross_entropy)
    V, h, b, s = 8192 + 200, 16, 2, 24

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(V, h)
            self.poison = False

        def forward(self, tokens, position_ids, attention_mask, txt_bool, img_bool, is_sparse, *mems):
            assert txt_bool.shape == img_bool.shape == tokens.shape and attention_mask.shape[-1] == tokens.shape[1]
            x = self.emb(tokens)
            logits = x @ self.emb.weight.t()
            if self.poison:
                logits = logits * float("nan")
            return (logits,)
    torch.manual_seed(0)
    m = Tiny()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.5 ** it)
    g = torch.Generator().manual_seed(1)
    tokens_full = torch.cat((torch.randint(8192, V, (b, 8), generator=g), torch.randint(0, 8192, (b, s - 7), generator=g)), 1)
    from cogview_b200 import data
    batch = data.make_batch({"text": tokens_full.numpy(), "loss_mask": np.ones((b, s + 1))})
    w0 = m.emb.weight.detach().clone()
    loss, skipped, mems, img_loss, txt_loss = pretrain.train_step(batch, m, opt, sched, txt_loss_scale=0.5)
    assert skipped == 0 and torch.isfinite(loss) and mems == [] and img_loss > 0 and txt_loss > 0
    assert not torch.equal(m.emb.weight, w0) and sched.last_epoch == 1
    m.poison = True
    w1 = m.emb.weight.detach().clone()
    out, skipped, *_ = pretrain.train_step(batch, m, opt, sched, txt_loss_scale=0.5)
    assert skipped == 1 and torch.isnan(out) and torch.equal(m.emb.weight, w1) and sched.last_epoch == 1


def test_query_templates_build_and_split_round_trip():
    from cogview_b200 import generate as gen
    tok = _tok()
    text = [8192 + 10, 8192 + 11, 8192 + 12]
    seq = gen.build_query(gen.QUERY_TEMPLATES['text2image'], [text], tok)
    assert seq[:6] == [tok['[ROI1]']] + text + [tok['[BASE]'], tok['[BOI1]']] and seq[6:] == [-1] * 1024
    codes = list(range(1024))
    seq = gen.build_query(gen.QUERY_TEMPLATES['low-level super-resolution'], [text, codes], tok)
    assert seq.count(-1) == 1024 and len(seq) == 1 + 3 + 2 + 1024 + 5 + 1024
    assert seq[6:6 + 1024] == codes and seq[6 + 1024:6 + 1024 + 5] == [tok['[EOI1]'], tok['[ROI2]'], tok['[POS0]'],
                                                                   tok['[BASE]'], tok['[BOI2]']]
    half = gen.build_query('[BASE] [BOI1] [Image512]{}', [codes], tok)          # keep 512 codes, generate the rest
    assert half[2:514] == codes[:512] and half[514:] == [-1] * 512
    with pytest.raises(ValueError):
        gen.build_query('[ROI1] a_raw_word [BASE]', [], tok)
    filled = [c if c >= 0 else 7 for c in gen.build_query(gen.QUERY_TEMPLATES['post-selection'], [codes, text], tok)]
    parts, images = gen.split_tokens(filled, tok)
    assert images == [codes] and parts == ['[BASE]', '[BOI1]', '[EOI1]', '[ROI1]', [10, 11, 12]]


def test_generate_images_once_batches_beams_and_decodes_the_last_image():
    from cogview_b200 import generate as gen
    tok = _tok()
    seq = torch.tensor(gen.build_query(gen.QUERY_TEMPLATES['text2image'], [[8192 + 5]], tok))

    class Args:
        max_inference_batch_size = 4
    calls = []

    def fake_fill(model, s, args):
        nb = -int(s[s < 0][0])
        assert (s[s < 0] == -nb).all()
        calls.append(nb)
        rows = s.clone().unsqueeze(0).repeat(nb, 1)
        rows[rows < 0] = len(calls)                        # "generated" code = index of the call
        return rows

    def fake_decode(codes):
        assert codes.shape == (1, 1024)
        return codes.float().mean().view(1, 1, 1, 1).expand(1, 3, 256, 256)
    rows, imgs = gen.generate_images_once(torch.nn.Identity(), None, Args, seq, num=8, fill=fake_fill, decode=fake_decode)
    assert calls == [4, 4] and rows.shape == (8, len(seq)) and imgs.shape == (8, 3, 256, 256)
    assert imgs[:4].unique().tolist() == [1.0] and imgs[4:].unique().tolist() == [2.0]
    assert (seq < 0).sum() == 1024                                             # the caller's template is untouched
    rows, imgs = gen.generate_images_once(torch.nn.Identity(), None, Args, seq, num=2, fill=fake_fill, decode=fake_decode)
    assert calls[-1] == 2 and rows.shape[0] == 2
    with pytest.raises(AssertionError):
        gen.generate_images_once(torch.nn.Identity(), None, Args, seq, num=6, fill=fake_fill, decode=fake_decode)


def test_inverse_prompt_score_with_a_stand_in_model():
    """generation/sampling.py:214-230 (post-selection): sum over the caption of log P(text token | image, previous text)
    with the image vocabulary masked out of the softmax — checked against the formula written out token by token."""
    from cogview_b200.generation import sampling

    class Args:
        is_sparse = 0
        img_tokenizer_num_tokens = 8192

    tok = sampling.get_tokenizer(Args)
    vocab = tok.num_tokens
    g = torch.Generator().manual_seed(5)
    botext = 2 + 1024 + 1
    ncap = 9
    seqs = []
    for i in range(2):
        img = torch.randint(0, 8192, (1024,), generator=g).tolist()
        cap = torch.randint(8192, 58192, (ncap,), generator=g).tolist()
        seqs.append([tok['[BASE]'], tok['[BOI1]']] + img + [tok['[EOI1]'], tok['[ROI1]']] + cap)
    seq = torch.tensor(seqs, dtype=torch.long)
    table = 0.15 * torch.randn(vocab, 64, generator=g)     # small logits: the reference's log(softmax()) does not underflow

    class StandIn:
        """logits[b, t] depend on the token at t (and on t): enough to pin indexing, masking and the gather"""

        def __call__(self, tokens, position_ids, attention_mask, txt, img, is_sparse, *mems):
            assert txt is None and img is None and is_sparse == 0
            assert attention_mask.shape[-2:] == (tokens.shape[1], tokens.shape[1])
            h = table[tokens] + 0.01 * position_ids.unsqueeze(-1).float()
            return (h @ table.t(),)

    got = sampling.inverse_prompt_score(StandIn(), seq, Args)
    assert got.shape == (2,)
    for b in range(2):
        total = 0.0
        for t in range(botext, seq.shape[1] - 1):
            lg = (table[seq[b, t]] + 0.01 * t) @ table.t()
            lg[:8192] = -float('inf')
            total += torch.log_softmax(lg, -1)[seq[b, t + 1]].item()
        assert abs(got[b].item() - total) < 1e-3 * abs(total), (got[b].item(), total)
    with pytest.raises(AssertionError):
        sampling.inverse_prompt_score(StandIn(), seq[:, 1:], Args)      # [ROI1] not where the layout puts it
