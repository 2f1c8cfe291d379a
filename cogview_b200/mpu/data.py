"""broadcast_data (reference mpu/data.py:76-116): at model-parallel size 1 the source rank is the only
member of its group, so the call reduces to moving the named tensors to the current device."""
import torch


def broadcast_data(keys, data, datatype):
    out = {}
    for k in keys:
        t = data[k]
        assert t.dtype == datatype, '{} has data type {} which is different than {}'.format(k, t.dtype, datatype)
        out[k] = t.cuda(non_blocking=True) if torch.cuda.is_available() else t
    return out
