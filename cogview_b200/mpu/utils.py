"""Helpers kept from the reference's mpu/utils.py:20-80 (same names and behaviour)."""
import torch


def ensure_divisibility(numerator, denominator):
    assert numerator % denominator == 0, '{} is not divisible by {}'.format(numerator, denominator)


def divide(numerator, denominator):
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def split_tensor_along_last_dim(tensor, num_partitions, contiguous_split_chunks=False):
    last_dim_size = divide(tensor.size()[-1], num_partitions)
    chunks = torch.split(tensor, last_dim_size, dim=tensor.dim() - 1)
    if contiguous_split_chunks:
        return tuple(c.contiguous() for c in chunks)
    return chunks


class VocabUtility:
    """[first, last) vocabulary range of a partition (mpu/utils.py:54-69)."""

    @staticmethod
    def vocab_range_from_per_partition_vocab_size(per_partition_vocab_size, rank, world_size):
        first = rank * per_partition_vocab_size
        return first, first + per_partition_vocab_size

    @staticmethod
    def vocab_range_from_global_vocab_size(global_vocab_size, rank, world_size):
        return VocabUtility.vocab_range_from_per_partition_vocab_size(divide(global_vocab_size, world_size), rank,
                                                                      world_size)
