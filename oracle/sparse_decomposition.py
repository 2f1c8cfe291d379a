"""TEST INFRASTRUCTURE (CPU): the tile-schedulable form of the reference's sparse TRAINING attention
(mpu/sparse_transformer.py:675-725, row a7 of SURVEY §8), proven equal to the restatement in cogview_oracle.py.

The reference evaluates one softmax over [gathered pivot keys | a causal band of `key_window_times` windows].
Because masked entries carry -10000 (exp underflows to exactly 0 in fp32) this is the same as a softmax over the
UNION of two disjoint visible key sets, which a flash-style kernel can walk as two masked passes that share one
running (max, sum):

  band   : keys j with band_start(i) <= j <= i,          band_start(i) = max(0, i // w - times + 1) * w
  pivots : gathered keys p with pos[p] < band_start(i),  score + log(s // n_piv)     (rmask of :491-496, gathered :569)

and merged through the joint log-sum-exp.  The backward of each pass is the ordinary flash backward evaluated with
the JOINT lse and the JOINT delta = rowsum(dO * O); the pivot pass scatters dK / dV back to the pivot positions.
Nothing here is imported by the product."""
import math

import torch


def band_start(i, w, times):
    """First key of query i's causal band: the window of i and the (times - 1) windows before it."""
    return torch.clamp(i // w - times + 1, min=0) * w


def visibility(s, pivot_idx, w, times):
    """band [s, s] and pivot [b, s, n_piv] boolean visibility in closed form."""
    i = torch.arange(s).unsqueeze(1)
    j = torch.arange(s).unsqueeze(0)
    bs = band_start(i, w, times)
    band = (j >= bs) & (j <= i)
    piv = pivot_idx.unsqueeze(1) < bs.unsqueeze(0)          # [b, 1, n_piv] < [1, s, 1]
    return band, piv


def _partial_softmax(scores, visible):
    """Row-wise (lse, unnormalised-prob-normalised-by-own-lse) over the visible entries; rows that see nothing give
    lse = -inf and zero probabilities."""
    sc = scores.masked_fill(~visible, -float('inf'))
    lse = torch.logsumexp(sc, dim=-1, keepdim=True)
    p = torch.exp(sc - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse))
    return lse, torch.where(torch.isinf(lse), torch.zeros_like(p), p)


def sparse_attention_two_pass(q, k, v, pivot_idx, w, times):
    """q, k, v: [b, heads, s, hn]; pivot_idx: [b, n_piv].  Returns (out, joint lse [b, heads, s, 1])."""
    b, nh, s, hn = q.shape
    n_piv = pivot_idx.shape[1]
    band, piv = visibility(s, pivot_idx, w, times)
    qs = q / math.sqrt(hn)
    gidx = pivot_idx.view(b, 1, n_piv, 1).expand(b, nh, n_piv, hn)
    pk, pv = torch.gather(k, 2, gidx), torch.gather(v, 2, gidx)
    lse_b, p_b = _partial_softmax(torch.matmul(qs, k.transpose(-1, -2)), band)
    lse_p, p_p = _partial_softmax(torch.matmul(qs, pk.transpose(-1, -2)) + math.log(s // n_piv), piv.unsqueeze(1))
    lse = torch.logaddexp(lse_b, lse_p)
    out = torch.exp(lse_b - lse) * torch.matmul(p_b, v) + torch.exp(lse_p - lse) * torch.matmul(p_p, pv)
    return out, lse


def sparse_attention_two_pass_backward(q, k, v, pivot_idx, w, times, out, lse, d_out):
    """Flash-style backward of the two passes with the joint lse / delta.  Returns (dq, dk, dv)."""
    b, nh, s, hn = q.shape
    n_piv = pivot_idx.shape[1]
    scale = 1.0 / math.sqrt(hn)
    band, piv = visibility(s, pivot_idx, w, times)
    gidx = pivot_idx.view(b, 1, n_piv, 1).expand(b, nh, n_piv, hn)
    pk, pv = torch.gather(k, 2, gidx), torch.gather(v, 2, gidx)
    delta = (d_out * out).sum(-1, keepdim=True)
    dq = torch.zeros_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    # band pass
    p = torch.exp(torch.matmul(q * scale, k.transpose(-1, -2)) - lse).masked_fill(~band, 0.0)
    ds = p * (torch.matmul(d_out, v.transpose(-1, -2)) - delta)
    dq += torch.matmul(ds, k) * scale
    dk += torch.matmul(ds.transpose(-1, -2), q) * scale
    dv += torch.matmul(p.transpose(-1, -2), d_out)
    # pivot pass (bias shifts the scores, not the gradients)
    p = torch.exp(torch.matmul(q * scale, pk.transpose(-1, -2)) + math.log(s // n_piv) - lse)
    p = p.masked_fill(~piv.unsqueeze(1), 0.0)
    ds = p * (torch.matmul(d_out, pv.transpose(-1, -2)) - delta)
    dq += torch.matmul(ds, pk) * scale
    dk.scatter_add_(2, gidx, torch.matmul(ds.transpose(-1, -2), q) * scale)
    dv.scatter_add_(2, gidx, torch.matmul(p.transpose(-1, -2), d_out))
    return dq, dk, dv
