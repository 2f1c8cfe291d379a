"""Data-parallel wrappers — mirror of the reference's model/distributed.py: `PyTorchDistributedDataParallel`
(:26-32, torch DDP whose state_dict is the bare module's) and `DistributedDataParallel` (:35-101, the
flatten -> all-reduce -> unflatten wrapper driven by `allreduce_params`)."""
import torch
import torch.distributed as dist
from torch.nn.modules import Module
from torch.nn.parallel.distributed import DistributedDataParallel as DDP

from .. import mpu

__all__ = ['PyTorchDistributedDataParallel', 'DistributedDataParallel']


class PyTorchDistributedDataParallel(DDP):
    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)


def _average_flat(tensors, group, world, *, pre_scale, post_scale, in_fp32):
    """One collective for a list of same-dtype tensors: pack them into a single 1-D buffer, all-reduce it over `group`,
    divide by the group size before or after the sum (or not at all) and scatter the result back in place."""
    sizes = [t.numel() for t in tensors]
    flat = torch.cat([t.reshape(-1) for t in tensors])
    if in_fp32:
        flat = flat.float()
    if pre_scale:
        flat.div_(world)
    dist.all_reduce(flat, group=group)
    if post_scale:
        flat.div_(world)
    for dst, piece in zip(tensors, flat.split(sizes)):
        dst.copy_(piece.view_as(dst))


class DistributedDataParallel(Module):
    """Same surface as the reference's wrapper (model/distributed.py:35-101): parameters are broadcast from rank 0 at
    construction, `forward` marks the gradients as stale, and `allreduce_params(reduce_after=True, no_scale=False,
    fp32_allreduce=False)` averages every existing gradient over the data-parallel group with one all-reduce per dtype
    (NCCL over NVLink / NVSwitch on the GPU box, gloo on CPU).  The training bench uses the torch-DDP subclass above (bucketed,
    overlapped with the backward); this wrapper is the reference's explicit-call alternative."""

    def __init__(self, module):
        super().__init__()
        self.module = module
        self.data_parallel_group = mpu.get_data_parallel_group()
        for p in module.parameters():
            dist.broadcast(p.data, 0, group=self.data_parallel_group)
        self.needs_reduction = False

    def allreduce_params(self, reduce_after=True, no_scale=False, fp32_allreduce=False):
        if not self.needs_reduction:
            return
        self.needs_reduction = False
        group = self.data_parallel_group
        world = dist.get_world_size(group=group)
        by_dtype = {}
        for p in self.module.parameters():
            if p.requires_grad and p.grad is not None:
                by_dtype.setdefault(p.dtype, []).append(p.grad.data)
        for grads in by_dtype.values():
            _average_flat(grads, group, world, pre_scale=not no_scale and not reduce_after,
                          post_scale=not no_scale and reduce_after, in_fp32=fp32_allreduce)

    def forward(self, *inputs, **kwargs):
        self.needs_reduction = True
        return self.module(*inputs, **kwargs)

    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)
