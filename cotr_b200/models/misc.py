"""NestedTensor plumbing of the reference (COTR/models/misc.py:35-80), reduced to what the hot path needs.

In the reference the mask is all-False for the fixed 256x512 canvas (misc.py:73-77 + backbone.py:80), so the native
path never materialises it; the type is kept because callers may pass one.
"""
import torch


class NestedTensor(object):
    def __init__(self, tensors, mask=None):
        self.tensors = tensors
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), None if self.mask is None else self.mask.to(device))

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list):
    """Stack equally-sized CHW images into a batch with an all-False padding mask (misc.py:58-80)."""
    if isinstance(tensor_list, torch.Tensor) and tensor_list.ndim == 4:
        batch = tensor_list
    else:
        if tensor_list[0].ndim != 3:
            raise ValueError('not supported')
        shapes = {tuple(t.shape) for t in tensor_list}
        if len(shapes) != 1:
            raise ValueError('the native path only supports equally sized 3x256x512 canvases')
        batch = torch.stack(list(tensor_list))
    mask = torch.zeros((batch.shape[0],) + tuple(batch.shape[-2:]), dtype=torch.bool, device=batch.device)
    return NestedTensor(batch, mask)
