"""The four tensor-parallel region mappings (reference mpu/mappings.py:79-141).  With the model-parallel
degree fixed at 1 each is the identity in forward and backward, exactly what the reference's short-circuits
at :27, :42, :61 do for a size-1 group."""


def copy_to_model_parallel_region(input_):
    return input_


def reduce_from_model_parallel_region(input_):
    return input_


def scatter_to_model_parallel_region(input_):
    return input_


def gather_from_model_parallel_region(input_):
    return input_
