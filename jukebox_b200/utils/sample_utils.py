"""Window / batch splitting helpers with the reference's names and results
(jukebox/utils/sample_utils.py:4-22), used by sample.py to bound the batch per engine call and to lay the
sliding windows over a level."""
import torch as t


def _chunks(x, size):
    return t.split(x, size, dim=0)


def split_batch(obj, n_samples, split_size):
    """Split along dim 0 into pieces of at most `split_size`.
    Tensor -> tuple of tensors; list of tensors -> list of per-piece tuples; None -> one None per piece."""
    if obj is None:
        pieces = -(-n_samples // split_size)
        return [None for _ in range(pieces)]
    if isinstance(obj, t.Tensor):
        return _chunks(obj, split_size)
    if isinstance(obj, list):
        per_item = [_chunks(item, split_size) for item in obj]
        return [tuple(parts) for parts in zip(*per_item)]
    raise TypeError('Unknown input type')


def get_starts(total_length, n_ctx, hop_length):
    """Starts of the windows of length `n_ctx` that cover [0, total_length) with hop `hop_length`; a window that
    would run past the end is pulled back so that it ends exactly at the end."""
    last = total_length - n_ctx
    out = []
    start = 0
    while start < last + hop_length:
        out.append(min(start, last))
        start += hop_length
    return out
