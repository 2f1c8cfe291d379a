"""clip_grad_norm (reference mpu/grads.py:28-74) at model-parallel size 1: plain global-norm clipping."""
import torch

inf = float('inf')


def clip_grad_norm(parameters, max_norm, norm_type=2):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    parameters = [p for p in parameters if p.grad is not None]
    max_norm = float(max_norm)
    norm_type = float(norm_type)
    if len(parameters) == 0:
        return 0.0
    if norm_type == inf:
        total_norm = max(p.grad.detach().abs().max() for p in parameters)
        total_norm = float(total_norm)
    else:
        norms = torch._foreach_norm([p.grad.detach() for p in parameters], norm_type)
        total_norm = float(torch.linalg.vector_norm(torch.stack([n.float() for n in norms]), norm_type))
    clip_coef = max_norm / (total_norm + 1e-6)
    if clip_coef < 1:
        torch._foreach_mul_([p.grad for p in parameters], clip_coef)
    return total_norm
