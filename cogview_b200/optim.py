"""Fused AdamW for bf16 models with fp32 master weights — the B200 replacement of the reference's
FP16_Optimizer + apex FusedAdam + mpu.clip_grad_norm sequence (pretrain_gpt2.py:110-158, :380-389, :437-444;
fp16/fp16.py:291-310, :399-453; mpu/grads.py:28-74).  bf16 needs no loss scaling, so the dynamic loss scaler
disappears; the NaN/inf guard of train_step (pretrain_gpt2.py:415-417) stays with the caller.

The clip coefficient is computed on the device (cv_sumsq_bf16 + cv_clip_coef) and consumed by cv_adamw_step
through a device pointer, so a step never synchronises with the host."""
import torch

from ._lib import check, lib, ptr, stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        self._scal = None
        self.last_grad_norm = None   # device tensor (1,) after step() when clipping is on

    def _state_for(self, p):
        st = self.state[p]
        if not st:
            st['step'] = 0
            st['master'] = p.detach().float().clone()
            st['exp_avg'] = torch.zeros_like(st['master'])
            st['exp_avg_sq'] = torch.zeros_like(st['master'])
        return st

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        L = lib()
        stream = stream_ptr()
        plist = [(g, p) for g in self.param_groups for p in g['params'] if p.grad is not None]
        if not plist:
            return None
        dev = plist[0][1].device
        coef = None
        if self.max_grad_norm > 0:
            if self._scal is None or self._scal.device != dev:
                self._scal = torch.zeros(3, dtype=torch.float32, device=dev)
            self._scal.zero_()
            for _, p in plist:
                g = p.grad
                assert g.dtype == torch.bfloat16 and g.is_contiguous(), "FusedAdamW expects contiguous bf16 gradients"
                check(L.cv_sumsq_bf16(ptr(g), g.numel(), ptr(self._scal[0:1]), stream), "cv_sumsq_bf16")
            check(L.cv_clip_coef(ptr(self._scal[0:1]), self.max_grad_norm, ptr(self._scal[1:2]), ptr(self._scal[2:3]),
                                 stream), "cv_clip_coef")
            coef = self._scal[1:2]
            self.last_grad_norm = self._scal[2:3]
        for group, p in plist:
            assert p.dtype == torch.bfloat16 and p.is_contiguous(), "FusedAdamW expects contiguous bf16 parameters"
            st = self._state_for(p)
            st['step'] += 1
            b1, b2 = group['betas']
            check(L.cv_adamw_step(ptr(p), ptr(p.grad), ptr(st['master']), ptr(st['exp_avg']), ptr(st['exp_avg_sq']),
                                  p.numel(), float(group['lr']), float(b1), float(b2), float(group['eps']),
                                  float(group['weight_decay']), int(st['step']), ptr(coef), 1.0, stream),
                  "cv_adamw_step")
        return None
