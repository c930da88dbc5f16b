"""Window / batch splitting helpers (reference: jukebox/utils/sample_utils.py)."""
import torch as t


def split_batch(obj, n_samples, split_size):
    n_passes = (n_samples + split_size - 1) // split_size
    if isinstance(obj, t.Tensor):
        return t.split(obj, split_size, dim=0)
    if isinstance(obj, list):
        return list(zip(*[t.split(item, split_size, dim=0) for item in obj]))
    if obj is None:
        return [None] * n_passes
    raise TypeError('Unknown input type')


def get_starts(total_length, n_ctx, hop_length):
    """window starts covering total_length; the last window is pulled back to end exactly at the end"""
    starts = []
    for start in range(0, total_length - n_ctx + hop_length, hop_length):
        starts.append(total_length - n_ctx if start + n_ctx >= total_length else start)
    return starts
