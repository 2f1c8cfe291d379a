"""Process groups — mirror of the reference's mpu/initialize.py:30-127 with the model-parallel degree fixed
at 1 (north star): the data-parallel group is the world, the model-parallel group is this rank alone."""
import torch

_MODEL_PARALLEL_GROUP = None
_DATA_PARALLEL_GROUP = None
_INITIALIZED = False


def initialize_model_parallel(model_parallel_size_=1):
    """mpu/initialize.py:30-78.  Only model_parallel_size 1 is supported (pure data parallelism)."""
    global _MODEL_PARALLEL_GROUP, _DATA_PARALLEL_GROUP, _INITIALIZED
    if int(model_parallel_size_) != 1:
        raise NotImplementedError("cogview_b200 fixes the model-parallel degree at 1 (data parallel only); got %r"
                                  % (model_parallel_size_,))
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        if torch.distributed.get_rank() == 0:
            print('> initializing model parallel with size 1')
        rank = torch.distributed.get_rank()
        world = torch.distributed.get_world_size()
        _DATA_PARALLEL_GROUP = torch.distributed.group.WORLD
        # one single-rank group per rank (every rank must take part in every new_group call)
        for r in range(world):
            g = torch.distributed.new_group([r])
            if r == rank:
                _MODEL_PARALLEL_GROUP = g
    _INITIALIZED = True


def model_parallel_is_initialized():
    return _INITIALIZED


def get_model_parallel_group():
    assert _INITIALIZED, 'model parallel group is not initialized'
    return _MODEL_PARALLEL_GROUP


def get_data_parallel_group():
    assert _INITIALIZED, 'data parallel group is not initialized'
    return _DATA_PARALLEL_GROUP


def get_model_parallel_world_size():
    return 1


def get_model_parallel_rank():
    return 0


def get_model_parallel_src_rank():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def get_data_parallel_world_size():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


def get_data_parallel_rank():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def destroy_model_parallel():
    global _MODEL_PARALLEL_GROUP, _DATA_PARALLEL_GROUP, _INITIALIZED
    _MODEL_PARALLEL_GROUP = None
    _DATA_PARALLEL_GROUP = None
    _INITIALIZED = False
