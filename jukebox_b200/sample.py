"""Windowed multi-level sampling (reference: jukebox/sample.py:17-147).

Same functions and argument meaning: sample_partial_window, sample_single_window, sample_level,
_sample, ancestral_sample, continue_sample, upsample, primed_sample.  Wav / HTML / alignment output
(reference :110-120) is file I/O and out of scope: `_sample` returns the codes and, when
hps.get('save_dir') is set, writes the reference's `data.pth.tar` resume format per level."""
import os

import torch as t

from .utils import dist_adapter as dist
from .utils.dist_adapter import print_once
from .utils.torch_utils import empty_cache
from .utils.sample_utils import split_batch, get_starts


def sample_partial_window(zs, labels, sampling_kwargs, level, prior, tokens_to_sample, hps):
    """a window shorter than n_ctx: `tokens_to_sample` new tokens at `level`"""
    z = zs[level]
    n_ctx = prior.n_ctx
    current_tokens = z.shape[1]
    if current_tokens < n_ctx - tokens_to_sample:
        sampling_kwargs['sample_tokens'] = current_tokens + tokens_to_sample
        start = 0
    else:
        sampling_kwargs['sample_tokens'] = n_ctx
        start = current_tokens - n_ctx + tokens_to_sample
    return sample_single_window(zs, labels, sampling_kwargs, level, prior, start, hps)


def sample_single_window(zs, labels, sampling_kwargs, level, prior, start, hps):
    """one window of n_ctx tokens starting at `start`; already-sampled tokens are the prime"""
    n_samples = hps.n_samples
    n_ctx = prior.n_ctx
    end = start + n_ctx
    z = zs[level][:, start:end]
    sample_tokens = sampling_kwargs.get('sample_tokens', end - start)
    conditioning_tokens, new_tokens = z.shape[1], sample_tokens - z.shape[1]
    print_once(f"Sampling {sample_tokens} tokens for [{start},{start + sample_tokens}]. "
               f"Conditioning on {conditioning_tokens} tokens")
    if new_tokens <= 0:
        return zs
    z_conds = prior.get_z_conds(zs, start, end)
    y = prior.get_y(labels, start)
    kwargs = dict(sampling_kwargs)
    max_batch_size = kwargs.pop('max_batch_size')
    z_list = split_batch(z, n_samples, max_batch_size)
    z_conds_list = split_batch(z_conds, n_samples, max_batch_size)
    y_list = split_batch(y, n_samples, max_batch_size)
    z_samples = []
    for z_i, z_conds_i, y_i in zip(z_list, z_conds_list, y_list):
        z_conds_i = None if z_conds_i is None else [c.contiguous() for c in z_conds_i]
        z_samples.append(prior.sample(n_samples=z_i.shape[0], z=z_i, z_conds=z_conds_i, y=y_i, **kwargs))
    z = t.cat(z_samples, dim=0)
    z_new = z[:, -new_tokens:]
    zs[level] = t.cat([zs[level], z_new], dim=1)
    return zs


def sample_level(zs, labels, sampling_kwargs, level, prior, total_length, hop_length, hps):
    print_once(f"Sampling level {level}")
    if total_length >= prior.n_ctx:
        for start in get_starts(total_length, prior.n_ctx, hop_length):
            zs = sample_single_window(zs, labels, sampling_kwargs, level, prior, start, hps)
    else:
        zs = sample_partial_window(zs, labels, sampling_kwargs, level, prior, total_length, hps)
    return zs


def _sample(zs, labels, sampling_kwargs, priors, sample_levels, hps):
    xs = {}
    for level in reversed(sample_levels):
        prior = priors[level]
        prior.cuda()
        assert hps.sample_length % prior.raw_to_tokens == 0, \
            f"Expected sample_length {hps.sample_length} to be multiple of {prior.raw_to_tokens}"
        total_length = hps.sample_length // prior.raw_to_tokens
        hop_length = int(hps.hop_fraction[level] * prior.n_ctx)
        zs = sample_level(zs, labels[level], sampling_kwargs[level], level, prior, total_length, hop_length, hps)
        if hps.get('offload_priors', False):      # the reference always did (16 GB cards); 180 GB keeps them
            prior.cpu()
            empty_cache()
        x = prior.decode(zs[level:], start_level=level, bs_chunks=zs[level].shape[0])
        xs[level] = x
        save_dir = hps.get('save_dir', None)
        if save_dir:
            name = f"{save_dir}_rank_{dist.get_rank()}" if dist.get_world_size() > 1 else save_dir
            logdir = f"{name}/level_{level}"
            os.makedirs(logdir, exist_ok=True)
            t.save(dict(zs=zs, labels=labels, sampling_kwargs=sampling_kwargs, x=x), f"{logdir}/data.pth.tar")
    hps['_last_audio'] = xs
    return zs


def ancestral_sample(labels, sampling_kwargs, priors, hps):
    sample_levels = list(range(len(priors)))
    dev = 'cuda'
    zs = [t.zeros(hps.n_samples, 0, dtype=t.long, device=dev) for _ in range(len(priors))]
    return _sample(zs, labels, sampling_kwargs, priors, sample_levels, hps)


def continue_sample(zs, labels, sampling_kwargs, priors, hps):
    return _sample(zs, labels, sampling_kwargs, priors, list(range(len(priors))), hps)


def upsample(zs, labels, sampling_kwargs, priors, hps):
    return _sample(zs, labels, sampling_kwargs, priors, list(range(len(priors) - 1)), hps)


def primed_sample(x, labels, sampling_kwargs, priors, hps):
    zs = priors[-1].encode(x, start_level=0, end_level=len(priors), bs_chunks=x.shape[0])
    return _sample(zs, labels, sampling_kwargs, priors, list(range(len(priors))), hps)


def load_codes(codes_file, duration, priors, hps):
    data = t.load(codes_file, map_location='cpu', weights_only=False)
    zs = [z.cuda() for z in data['zs']]
    assert zs[-1].shape[0] == hps.n_samples, f"Expected bs = {hps.n_samples}, got {zs[-1].shape[0]}"
    if duration is not None:
        top_raw_to_tokens = priors[-1].raw_to_tokens
        assert duration % top_raw_to_tokens == 0
        assert duration // top_raw_to_tokens <= zs[-1].shape[1]
        zs = [z[:, :duration // prior.raw_to_tokens] for z, prior in zip(zs, priors)]
    return zs
