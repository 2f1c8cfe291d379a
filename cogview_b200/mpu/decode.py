"""Single-token decode step on the K|V cache, captured once as a CUDA graph and replayed per token.

This is the fast path behind GPT2Model.forward when it is called the way generation/sampling.py:147-151 calls
it (one new token per sequence, `mems` returned by the previous call, mems_mode 'kv').  Per layer:
abs-max LN -> QKV linear -> cached attention (+ append) -> out-proj (+abs-max) -> LN + residual -> LN ->
h->4h (+GELU) -> 4h->h (+abs-max) -> LN + residual, all weight-streaming kernels (cv_linear_small_m,
cv_attn_decode) with the position read from device memory so the graph is replayable."""
import torch
import torch.nn.functional as F

from .. import ops
from .layers import _as_bf16


class DecodeRunner:
    def __init__(self, model, caches, use_graph=True):
        tr = model.transformer
        self.model = model
        self.caches = caches
        self.b = caches.b
        self.heads = tr.num_attention_heads
        self.h = tr.hidden_size
        dev = caches.buf.device
        self.ids = torch.zeros((self.b, 1), dtype=torch.int64, device=dev)
        self.pos = torch.zeros((self.b, 1), dtype=torch.int64, device=dev)
        self.cur_len = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = None
        self.graph = None
        self.use_graph = use_graph
        # enough (batch, head, split) CTAs to cover the SMs
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # the key range of every (batch, head) is split so that ~8 CTAs per SM stream the K|V cache: the kernel is
        # latency-bound per CTA (4B, b=4: 160 (batch, head) pairs alone left it at ~0.7 us per 16 keys)
        self.nsplit = max(1, min(16, -(-8 * sms // (self.b * self.heads))))
        self.params = None
        self.graph_launches = 0   # kernels per captured step
        self.replays = 0
        # token-generation runs (model step + sampling in one graph, next token fed back on the device)
        self.sample_graphs = {}
        self.out_buf = torch.zeros((self.b, caches.maxlen + 1), dtype=torch.int64, device=dev)
        self.stepc = torch.zeros((1, 1), dtype=torch.int64, device=dev)
        self.score_acc = torch.zeros(self.b, dtype=torch.float32, device=dev)

    def _gather_params(self):
        tr = self.model.transformer
        self.params = [tuple(_as_bf16(p.detach()) for p in layer.param_list()) for layer in tr.layers]
        self.wte = _as_bf16(self.model.word_embeddings.weight.detach()).contiguous()
        self.wpe = _as_bf16(tr.position_embeddings.weight.detach()).contiguous()
        self.fl = (_as_bf16(tr.final_layernorm.weight.detach()), _as_bf16(tr.final_layernorm.bias.detach()),
                   tr.final_layernorm.eps)
        self.eps = tr.layers[0].layernorm_epsilon

    def _run(self):
        b, h, heads = self.b, self.h, self.heads
        L = len(self.params)
        scal = ops.new_scalars(2 * L + 1, self.ids.device)
        x = ops.embed_fwd(self.ids, self.pos, self.wte, self.wpe, scal[2 * L:2 * L + 1])
        prev_gemm = prev_am = prev_post = None
        for i, P in enumerate(self.params):
            (g1, b1, wqkv, bqkv, wd, bd, g3, b3, g2, b2, w1, bb1, w2, bb2, g4, b4) = P
            s = scal[2 * i:2 * i + 2]
            # x_i = x_{i-1} + LN4(mlp_out_{i-1});  xn = LN1(x_i)
            y, xn = ops.ln_pair_small_m(x, prev_gemm, prev_am, prev_post, (g1, b1), self.eps,
                                        want_res_out=prev_gemm is not None)
            if y is not None:
                x = y
            qkv = ops.linear_small_m(xn, wqkv, bqkv)
            ctx = ops.attn_decode(qkv, self.caches.buf[i], heads, cur_len_dev=self.cur_len, nsplit=self.nsplit)
            attn_out = ops.linear_small_m(ctx, wd, bd, absmax=s[0:1])
            # y = x + LN3(attn_out);  xn2 = LN2(y)
            x, xn2 = ops.ln_pair_small_m(x, attn_out, s[0:1], (g3, b3), (g2, b2), self.eps)
            h4 = ops.linear_small_m(xn2, w1, bb1, act=ops.ACT_GELU)
            prev_gemm = ops.linear_small_m(h4, w2, bb2, absmax=s[1:2])
            prev_am, prev_post = s[1:2], (g4, b4)
        _, xf = ops.ln_pair_small_m(x, prev_gemm, prev_am, prev_post, (self.fl[0], self.fl[1]), self.fl[2],
                                    want_res_out=False)
        self.logits = ops.linear_small_m(xf, self.wte, out_dtype=torch.float32)

    def step(self, ids, pos, t):
        """ids, pos: [b, 1] int64; t: tokens already cached.  Returns logits [b, V] fp32 (a static buffer)."""
        if self.params is None:
            self._gather_params()
        self.ids.copy_(ids)
        self.pos.copy_(pos)
        self.cur_len.fill_(t)
        if not self.use_graph:
            self._run()
        elif self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run()
                self._run()
            torch.cuda.current_stream().wait_stream(side)
            from .._lib import lib
            before = lib().cv_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
            self.graph_launches = int(lib().cv_launch_count() - before)
            self.graph.replay()
            self.replays += 1
        else:
            self.graph.replay()
            self.replays += 1
        return self.logits

    # -- generation runs: generation/sampling.py:147-183 with nothing left on the host per token -----------------
    def _sample_body(self, key):
        """One token: decode step, then the reference's sampling tail — temperature, invalid vocabulary slices,
        top-k (generation/sampling.py:24-33 with the boolean index_put written as the equivalent masked_fill so that
        it can be captured), softmax, torch.multinomial, beam log-probability — and the hand-over of the sampled
        token to the next step, all through device-side state."""
        temperature, top_k, inv = key
        self._run()
        logits = self.logits
        logits.div_(temperature)
        for a, z in inv:
            logits[:, a:z] = -float('Inf')
        if top_k > 0:
            kth = torch.topk(logits, top_k)[0][..., -1, None]
            logits.masked_fill_(logits < kth, -float('Inf'))
        probs = F.softmax(logits, dim=-1)
        prev = torch.multinomial(probs, num_samples=1)
        self.score_acc.add_(torch.log(torch.gather(probs, 1, prev)[:, 0]))
        self.out_buf.scatter_(1, self.stepc.expand(self.b, 1), prev)
        self.stepc.add_(1)
        self.ids.copy_(prev)
        self.pos.add_(1)
        self.cur_len.add_(1)

    def _reset_run(self, ids, pos, t):
        self.ids.copy_(ids)
        self.pos.copy_(pos)
        self.cur_len.fill_(t)
        self.stepc.zero_()
        self.score_acc.zero_()

    def sample_run(self, ids, pos, t, n_steps, temperature, top_k, invalid_slices):
        """n_steps tokens starting from `ids` ([b, 1], at positions `pos`, t tokens cached).  One graph replay per
        token.  Returns (tokens [b, n_steps] int64, summed log-probabilities [b] fp32)."""
        if self.params is None:
            self._gather_params()
        vocab = self.model.word_embeddings.weight.shape[0]
        key = (float(temperature), int(top_k), tuple(sl.indices(vocab)[:2] for sl in invalid_slices))
        assert n_steps <= self.out_buf.shape[1]
        self._reset_run(ids, pos, t)
        graph = self.sample_graphs.get(key)
        if graph is None and self.use_graph:
            dev = self.ids.device
            rng = torch.cuda.get_rng_state(dev)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):           # warm-up (allocator, lazy init); its side effects are undone below
                self._sample_body(key)
                self._sample_body(key)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.set_rng_state(rng, dev)
            self._reset_run(ids, pos, t)
            from .._lib import lib
            before = lib().cv_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._sample_body(key)
            self.graph_launches = int(lib().cv_launch_count() - before)   # this library's kernels per replayed token
            self.sample_graphs[key] = graph
        for _ in range(n_steps):
            if graph is None:
                self._sample_body(key)
            else:
                graph.replay()
        self.replays += n_steps
        return self.out_buf[:, :n_steps].clone(), self.score_acc.clone()
