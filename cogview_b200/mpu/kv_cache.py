"""K|V-cache form of the reference's `mems` (mpu/sparse_transformer.py:526-529, :615-626).

The reference caches per-layer HIDDEN STATES and re-normalises / re-projects the whole memory every step
(:320, :136-141).  In 'kv' mode the list returned by GPT2Model.forward has the same length (num_layers + 1)
and the same batch / time axes, but entry i (< num_layers) is a view [b, t, 2h] of layer i's key|value cache
and the last entry is an empty placeholder; `filling_sequence` only ever expands / slices the batch axis and
passes the list back, which both forms support.  The cache buffers are pre-allocated to max_memory_length
and appended in place (positions >= t only, so an older view stays valid).
"""
import torch


class _Caches:
    def __init__(self, owner, b, device):
        self.h = owner.hidden_size
        self.L = len(owner.layers)
        self.maxlen = owner.max_memory_length
        self.b = b
        self.t = 0
        self.buf = torch.empty((self.L, b, self.maxlen, 2 * self.h), dtype=torch.bfloat16, device=device)
        self.dummy = torch.empty((b, self.maxlen, 0), dtype=torch.bfloat16, device=device)
        self.sq = 0

    def matches(self, mems, b):
        """True if `mems` are exactly this cache's current views (so we can append in place)."""
        if len(mems) != self.L + 1 or b != self.b:
            return False
        m0 = mems[0]
        return (m0.data_ptr() == self.buf[0].data_ptr() and m0.shape[0] == b and m0.shape[2] == 2 * self.h
                and m0.stride(0) == self.buf.stride(1) and m0.stride(1) == self.buf.stride(2))

    def load_from(self, mems):
        t = mems[0].size(1)
        for i in range(self.L):
            self.buf[i, :, :t].copy_(mems[i])
        self.t = t

    def begin(self, sq):
        if self.t + sq > self.maxlen:
            raise NotImplementedError('KV-cache overflow: memory %d + %d new tokens > max_memory_length %d '
                                      '(sliding memories need mems_mode="hidden")' % (self.t, sq, self.maxlen))
        self.sq = sq

    def appender(self, i):
        t, sq, h = self.t, self.sq, self.h
        layer_buf = self.buf[i]

        def kv(k_new, v_new):
            layer_buf[:, t:t + sq, :h].copy_(k_new)
            layer_buf[:, t:t + sq, h:].copy_(v_new)
            cur = layer_buf[:, :t + sq]
            return cur[..., :h], cur[..., h:]
        return kv

    def views(self):
        self.t += self.sq
        out = [self.buf[i, :, :self.t] for i in range(self.L)]
        out.append(self.dummy[:, :self.t])
        return out


def prepare(owner, mems, b, sq):
    """Find (or build) the cache that `mems` refers to and get it ready for sq more tokens.  One cache (and its
    captured decode graph) is kept per batch size and reused across sequences."""
    dev = owner.position_embeddings.weight.device
    pool = owner.__dict__.setdefault('_kv_pool', {})
    c = pool.get(b)
    if c is not None and (c.maxlen != owner.max_memory_length or c.buf.device != dev):
        c = None
    if not mems:
        if c is None:
            c = _Caches(owner, b, dev)
        c.t = 0
    elif c is not None and c.matches(mems, b):
        c.t = mems[0].size(1)
    else:
        # beams were expanded / selected (generation/sampling.py:168-172, :188-198): copy into this batch size's cache
        if c is None:
            c = _Caches(owner, b, dev)
        c.load_from(mems)
    pool[b] = c
    owner._kv = c
    c.begin(sq)
    return c
