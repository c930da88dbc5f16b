import sys

from . import dist_adapter as dist


def get_range(x):
    """tqdm on rank 0 (reference utils/logger.py:8-15); plain iterator if tqdm is missing"""
    if dist.get_rank() == 0:
        try:
            from tqdm import tqdm
            return tqdm(x, leave=True, file=sys.stdout,
                        bar_format="{n_fmt}/{total_fmt} [{elapsed}<{remaining}, {rate_fmt}{postfix}]")
        except ImportError:
            return x
    return x
