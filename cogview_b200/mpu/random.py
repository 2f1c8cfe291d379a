"""CUDA RNG tracker and activation checkpointing — same API as the reference's mpu/random.py
(CudaRNGStatesTracker :119-186, model_parallel_cuda_manual_seed :198-233, checkpoint :273-378).

The tracker forks a named RNG stream around model-parallel-region dropout; checkpoint(fn, *args) re-runs
`fn` in the backward with the RNG streams restored so dropout masks repeat.  On CPU-only hosts the tracker
degrades to the CPU generator so that host-side logic stays testable without a GPU."""
import contextlib

import torch
from torch.utils.checkpoint import detach_variable

from .initialize import get_data_parallel_rank, get_model_parallel_rank

_MODEL_PARALLEL_RNG_TRACKER_NAME = 'model-parallel-rng'
_PARTITION_ACTIVATIONS = False

# Dropout inside the fused kernels is counter-based: a site is (seed, site id); the keep mask is a pure function of
# (seed, site id, element index), so forward, backward and a checkpoint recomputation regenerate identical masks.
# The seed follows torch's default generator (set by set_random_seed -> torch.manual_seed); site ids count up.
_DROPOUT_SITE = [0]


def next_dropout_site():
    """(seed, site id) for the next dropout site of this process."""
    _DROPOUT_SITE[0] = (_DROPOUT_SITE[0] + 1) & 0xFFFFFFFF
    return int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF, _DROPOUT_SITE[0]


def get_dropout_site_counter():
    return _DROPOUT_SITE[0]


def set_dropout_site_counter(v):
    _DROPOUT_SITE[0] = int(v)


def _get_state():
    return torch.cuda.get_rng_state() if torch.cuda.is_available() else torch.get_rng_state()


def _set_state(state):
    if torch.cuda.is_available():
        torch.cuda.set_rng_state(state)
    else:
        torch.set_rng_state(state)


def _seed(seed):
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


class CudaRNGStatesTracker:
    """Named RNG states; `fork(name)` runs a block under that state and stores the advanced state back."""

    def __init__(self):
        self.states_ = {}
        self.seeds_ = set()

    def reset(self):
        self.states_ = {}
        self.seeds_ = set()

    def get_states(self):
        return dict(self.states_)

    def set_states(self, states):
        self.states_ = states

    def add(self, name, seed):
        if seed in self.seeds_:
            raise Exception('seed {} already exists'.format(seed))
        self.seeds_.add(seed)
        if name in self.states_:
            raise Exception('cuda rng state {} already exists'.format(name))
        orig = _get_state()
        _seed(seed)
        self.states_[name] = _get_state()
        _set_state(orig)

    @contextlib.contextmanager
    def fork(self, name=_MODEL_PARALLEL_RNG_TRACKER_NAME):
        if name not in self.states_:
            raise Exception('cuda rng state {} is not added'.format(name))
        orig = _get_state()
        _set_state(self.states_[name])
        try:
            yield
        finally:
            self.states_[name] = _get_state()
            _set_state(orig)


_CUDA_RNG_STATE_TRACKER = CudaRNGStatesTracker()


def get_cuda_rng_tracker():
    return _CUDA_RNG_STATE_TRACKER


def model_parallel_cuda_manual_seed(seed):
    """Default stream <- `seed` (same across a model-parallel group), tracked stream <- seed + 2718 + mp_rank
    (mpu/random.py:198-233)."""
    offset = seed + 2718
    model_parallel_seed = offset + get_model_parallel_rank()
    if (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0:
        print('> initializing model parallel cuda seeds on global rank {}, model parallel rank {}, and data '
              'parallel rank {} with model parallel seed: {} and data parallel seed: {}'.format(
                  torch.distributed.get_rank() if torch.distributed.is_initialized() else 0,
                  get_model_parallel_rank(), get_data_parallel_rank(), model_parallel_seed, seed), flush=True)
    _CUDA_RNG_STATE_TRACKER.reset()
    _seed(seed)
    _CUDA_RNG_STATE_TRACKER.add(_MODEL_PARALLEL_RNG_TRACKER_NAME, model_parallel_seed)


def partition_activations_in_checkpoint(partition_activation):
    """Kept for API compatibility (mpu/random.py:48-52); partitioning needs model parallelism > 1."""
    global _PARTITION_ACTIVATIONS
    _PARTITION_ACTIVATIONS = bool(partition_activation) and False


class CheckpointFunction(torch.autograd.Function):
    """Run `run_function` without saving intermediates; recompute it in backward under the saved RNG states."""

    @staticmethod
    def forward(ctx, run_function, *args):
        ctx.run_function = run_function
        ctx.cpu_rng = torch.get_rng_state()
        ctx.dev_rng = _get_state()
        ctx.tracker_states = get_cuda_rng_tracker().get_states()
        ctx.dropout_site = get_dropout_site_counter()
        ctx.tensor_idx = [i for i, a in enumerate(args) if torch.is_tensor(a)]
        ctx.other = {i: a for i, a in enumerate(args) if not torch.is_tensor(a)}
        ctx.nargs = len(args)
        ctx.save_for_backward(*[args[i] for i in ctx.tensor_idx])
        with torch.no_grad():
            outputs = run_function(*args)
        return outputs

    @staticmethod
    def backward(ctx, *grads):
        if not torch.autograd._is_checkpoint_valid():
            raise RuntimeError("Checkpointing is not compatible with .grad(), please use .backward() if possible")
        saved = ctx.saved_tensors
        args = [None] * ctx.nargs
        for i, t in zip(ctx.tensor_idx, saved):
            args[i] = t
        for i, a in ctx.other.items():
            args[i] = a
        bwd_cpu, bwd_dev = torch.get_rng_state(), _get_state()
        bwd_tracker = get_cuda_rng_tracker().get_states()
        bwd_site = get_dropout_site_counter()
        set_dropout_site_counter(ctx.dropout_site)
        torch.set_rng_state(ctx.cpu_rng)
        _set_state(ctx.dev_rng)
        get_cuda_rng_tracker().set_states(ctx.tracker_states)
        detached = detach_variable(tuple(args))
        with torch.enable_grad():
            outputs = ctx.run_function(*detached)
        torch.set_rng_state(bwd_cpu)
        _set_state(bwd_dev)
        get_cuda_rng_tracker().set_states(bwd_tracker)
        set_dropout_site_counter(bwd_site)
        if torch.is_tensor(outputs):
            outputs = (outputs,)
        pairs = [(o, g) for o, g in zip(outputs, grads) if torch.is_tensor(o) and o.requires_grad and g is not None]
        torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])
        return (None,) + tuple(a.grad if torch.is_tensor(a) else None for a in detached)


def checkpoint(function, *args):
    """mpu/random.py:375-378."""
    return CheckpointFunction.apply(function, *args)
