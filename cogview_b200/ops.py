"""Thin functional wrappers over the C ABI: allocate outputs with torch, pass raw pointers + stream.

These are the only places that call into libcogview_b200.so; the `mpu` / `model` / `vqvae` mirrors of the
reference interface are built on them.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

ACT_NONE = 0
ACT_GELU = 1


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, bias=None, act=ACT_NONE, out_dtype=torch.bfloat16,
         absmax=None, want_preact=False, out=None, block_n=0):
    """C[M,N] = op(A)[M,K] @ op(B)[N,K]^T (+bias) (+GELU).

    a: [M,K] (or [K,M] when a_mn_major); b: [N,K] (or [K,N] when b_mn_major); both bf16, last dim contiguous.
    absmax: optional 1-element fp32 tensor updated with atomic max |C|.
    Returns C, or (C, preact) when want_preact.
    """
    require_cuda(a, b, bias, absmax)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, "inner dimensions differ: %d vs %d" % (K, Kb)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1
    pre = torch.empty((M, N), dtype=torch.bfloat16, device=a.device) if want_preact else None
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.numel() == N
    rc = lib().cv_gemm_bf16(ptr(a), int(a_mn_major), a.stride(0), ptr(b), int(b_mn_major), b.stride(0),
                            ptr(out), int(out.dtype == torch.float32), out.stride(0), ptr(pre), ptr(bias), int(act),
                            ptr(absmax), M, N, K, block_n, stream_ptr())
    check(rc, "cv_gemm_bf16")
    return (out, pre) if want_preact else out
