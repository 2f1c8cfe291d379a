"""Data formats feeding the training path: the compact binary token files and the
text+code template of the reference.

  BinaryDataset              data_utils/datasets.py:63-81      [N, 64 + 1024] int32 rows, memory-mapped
  compact_binary_process_fn  data_utils/datasets.py:119-128    row -> {'text', 'loss_mask'}
  TextCodeTemplate           data_utils/templates.py:52-65     [ROI1] text [BASE] [BOI1] code [EOI1]
  pad_to_len                 data_utils/datasets.py:91-101
  make_batch                 pretrain_gpt2.py:256-289          tokens / labels / loss mask / attention mask / positions

Host-side numpy / torch only; LMDB datasets, sentencepiece and the VQ-VAE image tokenizer front-end stay out of
scope (synthetic or pre-tokenised rows)."""
import numpy as np
import torch

from .generation.sampling import get_masks_and_position_ids, get_tokenizer


def TextCodeTemplate(text, code, tokenizer=None):
    """data_utils/templates.py:52-65 for already-tokenised text (numpy array or list of ids)."""
    tok = tokenizer or get_tokenizer()
    if isinstance(text, str):
        raise NotImplementedError('raw strings need the sentencepiece tokenizer, which is outside this package')
    code = tok.wrap_code(code)
    if isinstance(text, list) and isinstance(code, list):
        return [tok['[ROI1]']] + text + code
    return np.concatenate((np.array([tok['[ROI1]']]), np.asarray(text), np.asarray(code)), axis=0)


def pad_to_len(ret, max_len, tokenizer=None):
    """data_utils/datasets.py:91-101: right-pad with [PAD] (or truncate) to max_len; returns (row, valid length)."""
    tok = tokenizer or get_tokenizer()
    if len(ret) < max_len:
        return np.concatenate((ret, np.array([tok['[PAD]']] * (max_len - len(ret)))), axis=0), len(ret)
    return ret[:max_len], max_len


def compact_binary_process_fn(max_len, tokenizer=None):
    """data_utils/datasets.py:119-128: 64 text slots (-1 = unused) + 1024 image codes per row."""
    def process_fn(row):
        text, code = row[:64].astype(np.int64), row[64:].astype(np.int64)
        text = text[text > -1]
        ret, sep = pad_to_len(TextCodeTemplate(text, code, tokenizer), max_len, tokenizer)
        return {'text': ret, 'loss_mask': np.array([1] * sep + [0] * (len(ret) - sep))}
    return process_fn


def tokenized_process_fn(max_len, tokenizer=None):
    """data_utils/datasets.py:103-109: rows that were tokenised when saved."""
    def process_fn(row):
        ret, sep = pad_to_len(row.flatten(), max_len, tokenizer)
        return {'text': ret, 'loss_mask': np.array([1] * sep + [0] * (len(ret) - sep))}
    return process_fn


class BinaryDataset(torch.utils.data.Dataset):
    """data_utils/datasets.py:63-81."""

    def __init__(self, path, process_fn, length_per_sample=64 + 1024, dtype='int32', preload=False, **kwargs):
        assert length_per_sample is not None
        self.length_per_sample = length_per_sample
        self.dtype = np.dtype(dtype)
        self.process_fn = process_fn
        if preload:
            self.bin = np.fromfile(path, dtype=self.dtype).reshape(-1, length_per_sample)
        else:
            with open(path, 'rb') as fid:
                flen = fid.seek(0, 2) // self.dtype.itemsize
            self.bin = np.memmap(path, dtype=self.dtype, mode='r',
                                 shape=(flen // length_per_sample, length_per_sample))

    def __len__(self):
        return self.bin.shape[0]

    def __getitem__(self, index):
        return self.process_fn(self.bin[index])


def make_batch(data, device=None):
    """pretrain_gpt2.py:256-289 without the model-parallel broadcast (MP = 1): `data` holds 'text' and 'loss_mask'
    [b, s + 1].  Returns tokens, labels, loss_mask, attention_mask, position_ids."""
    tokens_ = torch.as_tensor(np.asarray(data['text'])).long()
    loss_mask = torch.as_tensor(np.asarray(data['loss_mask'])).float()
    if device is not None:
        tokens_, loss_mask = tokens_.to(device), loss_mask.to(device)
    labels = tokens_[:, 1:].contiguous()
    loss_mask = loss_mask[:, 1:].contiguous()
    tokens = tokens_[:, :-1].contiguous()
    attention_mask, loss_mask, position_ids = get_masks_and_position_ids(tokens, loss_mask=loss_mask)
    return tokens, labels, loss_mask, attention_mask, position_ids
