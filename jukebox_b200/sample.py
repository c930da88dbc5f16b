"""Windowed multi-level sampling on resident priors.

Entry points keep the reference's names and argument meaning (jukebox/sample.py:17-147) because notebooks and
scripts call them: sample_partial_window, sample_single_window, sample_level, _sample, ancestral_sample,
continue_sample, upsample, primed_sample, load_codes.  The work itself is organised differently:

  * `plan_windows` is a pure function: given how many tokens a level already has, how many it needs and the
    prior's context, it lists the (start, sample_tokens) windows to run.  It is what decides the stitching, so it
    is tested on the CPU against the reference's own loop (tests/test_sample_plan_cpu.py).
  * `LevelRun` owns one level's codes / labels / sampling options and executes windows: slice the context,
    fetch the per-window conditioning from the prior, split the batch into engine-sized pieces
    (`max_batch_size`), call `prior.sample`, append the new tokens.
  * priors stay on the GPU between levels (180 GB holds all three); `hps.offload_priors` restores the reference's
    cpu() shuffling.

Wav / HTML / alignment output (reference :110-120) is file I/O and out of scope: `_sample` returns the codes and,
when hps.get('save_dir') is set, writes the reference's `data.pth.tar` resume format per level."""
import os
from dataclasses import dataclass

import torch as t

from .utils import dist_adapter as dist
from .utils.dist_adapter import print_once
from .utils.torch_utils import empty_cache
from .utils.sample_utils import split_batch, get_starts


@dataclass(frozen=True)
class Window:
    start: int              # first token of the context handed to the prior
    sample_tokens: int      # length of that context once the window is done (<= n_ctx)


def plan_windows(have, total_length, n_ctx, hop_length):
    """Windows that extend a level holding `have` tokens to `total_length` tokens.

    total_length >= n_ctx: full windows at get_starts(total_length, n_ctx, hop_length), each filled up to n_ctx
    (windows that are already complete are still listed: running them is a no-op).
    total_length <  n_ctx: ONE window that adds `total_length` tokens to what is there, sliding the context so
    that it never exceeds n_ctx (the reference's sample_partial_window)."""
    if total_length >= n_ctx:
        return [Window(s, n_ctx) for s in get_starts(total_length, n_ctx, hop_length)]
    if have + total_length < n_ctx:
        return [Window(0, have + total_length)]
    return [Window(have + total_length - n_ctx, n_ctx)]


class LevelRun:
    """One level of one sampling job: the codes sampled so far and what is needed to extend them."""

    def __init__(self, zs, labels, sampling_kwargs, level, prior, hps):
        self.zs, self.labels, self.level, self.prior, self.hps = zs, labels, level, prior, hps
        opts = dict(sampling_kwargs)
        opts.pop('sample_tokens', None)           # per-window, set by run_window
        self.max_batch = opts.pop('max_batch_size')
        self.opts = opts

    def have(self):
        return self.zs[self.level].shape[1]

    def run_window(self, win):
        prior, level, n = self.prior, self.level, self.hps.n_samples
        context = self.zs[level][:, win.start:win.start + prior.n_ctx]
        given = context.shape[1]
        missing = win.sample_tokens - given
        print_once(f"Sampling {win.sample_tokens} tokens for [{win.start},{win.start + win.sample_tokens}]. "
                   f"Conditioning on {given} tokens")
        if missing <= 0:
            return
        upper = prior.get_z_conds(self.zs, win.start, win.start + prior.n_ctx)
        y = prior.get_y(self.labels, win.start)
        extra = {} if win.sample_tokens == prior.n_ctx else dict(sample_tokens=win.sample_tokens)
        pieces = zip(split_batch(context, n, self.max_batch), split_batch(upper, n, self.max_batch),
                     split_batch(y, n, self.max_batch))
        done = []
        for ctx_i, upper_i, y_i in pieces:
            if upper_i is not None:
                upper_i = [u.contiguous() for u in upper_i]
            done.append(prior.sample(n_samples=ctx_i.shape[0], z=ctx_i, z_conds=upper_i, y=y_i, **self.opts, **extra))
        fresh = t.cat(done, dim=0)[:, -missing:]
        self.zs[level] = t.cat([self.zs[level], fresh], dim=1)

    def extend_to(self, total_length, hop_length):
        for win in plan_windows(self.have(), total_length, self.prior.n_ctx, hop_length):
            self.run_window(win)
        return self.zs


# ---- the reference's entry points -------------------------------------------------------------------------
def sample_partial_window(zs, labels, sampling_kwargs, level, prior, tokens_to_sample, hps):
    """`tokens_to_sample` new tokens at `level`, the context sliding once it is full"""
    run = LevelRun(zs, labels, sampling_kwargs, level, prior, hps)
    have = run.have()
    if have + tokens_to_sample < prior.n_ctx:
        win = Window(0, have + tokens_to_sample)
    else:
        win = Window(have + tokens_to_sample - prior.n_ctx, prior.n_ctx)
    run.run_window(win)
    return zs


def sample_single_window(zs, labels, sampling_kwargs, level, prior, start, hps):
    """the window of prior.n_ctx tokens that starts at `start`; tokens already there are the prime"""
    run = LevelRun(zs, labels, sampling_kwargs, level, prior, hps)
    run.run_window(Window(start, sampling_kwargs.get('sample_tokens', prior.n_ctx)))
    return zs


def sample_level(zs, labels, sampling_kwargs, level, prior, total_length, hop_length, hps):
    print_once(f"Sampling level {level}")
    return LevelRun(zs, labels, sampling_kwargs, level, prior, hps).extend_to(total_length, hop_length)


def _on_gpu(module):
    return all(p.is_cuda for p in module.parameters())


def _sample(zs, labels, sampling_kwargs, priors, sample_levels, hps):
    audio = {}
    for level in sorted(sample_levels, reverse=True):        # coarsest level first
        prior = priors[level]
        if not _on_gpu(prior):                               # a resident prior keeps its packed decode engine
            prior.cuda()
        assert hps.sample_length % prior.raw_to_tokens == 0, \
            f"Expected sample_length {hps.sample_length} to be multiple of {prior.raw_to_tokens}"
        tokens_needed = hps.sample_length // prior.raw_to_tokens
        hop = int(hps.hop_fraction[level] * prior.n_ctx)
        zs = sample_level(zs, labels[level], sampling_kwargs[level], level, prior, tokens_needed, hop, hps)
        if hps.get('offload_priors', False):      # the reference always did (16 GB cards); 180 GB keeps them
            prior.cpu()
            empty_cache()
        audio[level] = prior.decode(zs[level:], start_level=level, bs_chunks=zs[level].shape[0])
        out_dir = hps.get('save_dir', None)
        if out_dir:
            save_level(out_dir, level, zs, labels, sampling_kwargs, audio[level])
    hps['_last_audio'] = audio
    return zs


def save_level(save_dir, level, zs, labels, sampling_kwargs, x):
    """the reference's resume file (sample.py:116): {zs, labels, sampling_kwargs, x} per level"""
    root = f"{save_dir}_rank_{dist.get_rank()}" if dist.get_world_size() > 1 else save_dir
    logdir = os.path.join(root, f"level_{level}")
    os.makedirs(logdir, exist_ok=True)
    t.save(dict(zs=zs, labels=labels, sampling_kwargs=sampling_kwargs, x=x), os.path.join(logdir, "data.pth.tar"))
    return logdir


def _all_levels(priors):
    return list(range(len(priors)))


def ancestral_sample(labels, sampling_kwargs, priors, hps):
    empty = [t.zeros(hps.n_samples, 0, dtype=t.long, device='cuda') for _ in priors]
    return _sample(empty, labels, sampling_kwargs, priors, _all_levels(priors), hps)


def continue_sample(zs, labels, sampling_kwargs, priors, hps):
    return _sample(zs, labels, sampling_kwargs, priors, _all_levels(priors), hps)


def upsample(zs, labels, sampling_kwargs, priors, hps):
    return _sample(zs, labels, sampling_kwargs, priors, _all_levels(priors)[:-1], hps)


def primed_sample(x, labels, sampling_kwargs, priors, hps):
    zs = priors[-1].encode(x, start_level=0, end_level=len(priors), bs_chunks=x.shape[0])
    return _sample(zs, labels, sampling_kwargs, priors, _all_levels(priors), hps)


def load_codes(codes_file, duration, priors, hps):
    """codes of a previous run (`data.pth.tar`), optionally cut to `duration` raw samples"""
    stored = t.load(codes_file, map_location='cpu', weights_only=False)['zs']
    assert stored[-1].shape[0] == hps.n_samples, f"Expected bs = {hps.n_samples}, got {stored[-1].shape[0]}"
    keep = [z.shape[1] for z in stored]
    if duration is not None:
        top = priors[-1].raw_to_tokens
        assert duration % top == 0, f"duration {duration} is not a multiple of {top}"
        assert duration // top <= stored[-1].shape[1]
        keep = [duration // prior.raw_to_tokens for prior in priors]
    return [z[:, :k].cuda() for z, k in zip(stored, keep)]
