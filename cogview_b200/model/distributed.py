"""Data-parallel wrappers — mirror of the reference's model/distributed.py: `PyTorchDistributedDataParallel`
(:26-32, torch DDP whose state_dict is the bare module's) and `DistributedDataParallel` (:35-101, the
flatten -> all-reduce -> unflatten wrapper driven by `allreduce_params`)."""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors
from torch.nn.modules import Module
from torch.nn.parallel.distributed import DistributedDataParallel as DDP

from .. import mpu

__all__ = ['PyTorchDistributedDataParallel', 'DistributedDataParallel']


class PyTorchDistributedDataParallel(DDP):
    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)


class DistributedDataParallel(Module):
    """Broadcast parameters from rank 0 at construction; `allreduce_params()` averages all gradients over the
    data-parallel group as one flat buffer per dtype (NCCL over NVLink / NVSwitch on the GPU box, gloo on CPU)."""

    def __init__(self, module):
        super().__init__()
        self.module = module
        self.data_parallel_group = mpu.get_data_parallel_group()
        src_rank = 0
        for p in self.module.parameters():
            dist.broadcast(p.data, src_rank, group=self.data_parallel_group)
        self.needs_reduction = False

        def allreduce_params(reduce_after=True, no_scale=False, fp32_allreduce=False):
            if not self.needs_reduction:
                return
            self.needs_reduction = False
            buckets = {}
            for _, param in self.module.named_parameters():
                if param.requires_grad and param.grad is not None:
                    buckets.setdefault(param.data.dtype, []).append(param)
            world = dist.get_world_size(group=self.data_parallel_group)
            for tp, bucket in buckets.items():
                grads = [p.grad.data for p in bucket]
                coalesced = _flatten_dense_tensors(grads)
                if fp32_allreduce:
                    coalesced = coalesced.float()
                if not no_scale and not reduce_after:
                    coalesced /= world
                dist.all_reduce(coalesced, group=self.data_parallel_group)
                if not no_scale and reduce_after:
                    coalesced /= world
                for buf, synced in zip(grads, _unflatten_dense_tensors(coalesced, grads)):
                    buf.copy_(synced)

        self.allreduce_params = allreduce_params

    def forward(self, *inputs, **kwargs):
        self.needs_reduction = True
        return self.module(*inputs, **kwargs)

    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)
