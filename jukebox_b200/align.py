"""Lyric alignment from attention weights (reference jukebox/align.py:15-84).

For every window ("hop") of the top-level codes the prior's forward pass is run with attention recording on the model's
designated alignment layer; the chosen head's [codes x lyric tokens] weights of each hop are scattered back to the columns
of the full lyric (the labeller tells which full-lyric index each of the window's n_tokens slots came from) and later hops
are overwritten by earlier ones, exactly as the reference's reversed loop does.

The forward pass is jukebox_b200's forward-mode path (Transformer.forward(sample=False) -> csrc/f32_path.cu).
"""
import numpy as np
import torch as t

from .utils.sample_utils import get_starts


def pad_to_context(z, n_ctx):
    """codes shorter than one context are right-padded with code 0; returns (z, pad)"""
    pad = max(0, n_ctx - z.shape[1])
    if pad:
        z = t.cat([z, t.zeros(z.shape[0], pad, dtype=z.dtype, device=z.device)], dim=1)
    return z, pad


def hop_weights(prior, z_window, y, fp16):
    """[bs, n_ctx, n_tokens] weights of the alignment head for one window; items go through one at a time like the
    reference (a recorded [1, heads, n_ctx, keys] tensor per item)"""
    layer, head = prior.alignment_layer, prior.alignment_head
    rows = []
    for i in range(z_window.shape[0]):
        ws = prior.z_forward(z_window[i:i + 1], [], y[i:i + 1], fp16=fp16, get_attn_weights={layer})
        assert len(ws) == 1
        rows.append(ws[0][:, head].float())
    w = t.cat(rows, dim=0)
    assert w.shape == (z_window.shape[0], prior.n_ctx, prior.n_tokens), tuple(w.shape)
    return w.cpu().numpy()


def stitch(hops, indices, starts, full_lengths, total_length, n_ctx, pad):
    """hops[start]: [bs, n_ctx, n_tokens]; indices[start][item]: full-lyric column of each token slot.
    -> per item [total_length - pad, len(full lyric)]"""
    out = []
    for item, n_full in enumerate(full_lengths):
        a = np.zeros((total_length, n_full + 1))
        for start in reversed(starts):
            a[start:start + n_ctx, indices[start][item]] = hops[start][item]
        out.append(a[:total_length - pad, :-1])      # drop the padding rows and the column of the "no token" slot
    return out


def get_alignment(x, zs, labels, prior, fp16, hps):
    """alignments: list (one per item) of [codes, lyric characters] attention maps - signature of the reference"""
    level = hps.levels - 1
    n_ctx, n_tokens = prior.n_ctx, prior.n_tokens
    z, pad = pad_to_context(zs[level], n_ctx)
    bs, total_length = z.shape
    hop = int(hps.hop_fraction[level] * n_ctx)
    starts = list(get_starts(total_length, n_ctx, hop))
    hops, indices = {}, {}
    with t.no_grad():
        for start in starts:
            y, idx = prior.get_y(labels, start, get_indices=True)
            assert len(idx) == bs and all(len(i) == n_tokens for i in idx)
            hops[start] = hop_weights(prior, z[:, start:start + n_ctx], y, fp16)
            indices[start] = idx
    full_lengths = [len(info['full_tokens']) for info in labels['info']]
    return stitch(hops, indices, starts, full_lengths, total_length, n_ctx, pad)
