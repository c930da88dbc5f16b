"""Host logic of lyric alignment (jukebox_b200/align.py; reference jukebox/align.py:15-84) with a stub prior: hop
placement, scatter through the labeller's indices, first-hop-wins overlap, padding of short pieces."""
import numpy as np
import torch

from jukebox_b200 import align
from jukebox_b200.utils.sample_utils import get_starts


class StubPrior:
    n_ctx, n_tokens, alignment_layer, alignment_head = 8, 3, 5, 1

    def __init__(self):
        self.calls = []

    def get_y(self, labels, start, get_indices=False):
        bs = len(labels['info'])
        y = torch.full((bs, 4), start)
        idx = [[(start // 4 + j) % (len(labels['info'][i]['full_tokens']) + 1) for j in range(self.n_tokens)]
               for i in range(bs)]
        return y, idx

    def z_forward(self, z, z_conds, y, fp16=False, get_attn_weights=False):
        assert get_attn_weights == {self.alignment_layer} and z.shape == (1, self.n_ctx) and z_conds == []
        self.calls.append(int(y[0, 0]))
        w = torch.zeros(1, 2, self.n_ctx, self.n_tokens)
        # head 1: code value + 0.25 * token slot, so the test can tell which hop / row / slot a value came from
        w[0, 1] = z[0].float()[:, None] + 0.25 * torch.arange(self.n_tokens)[None]
        return [w]


class Hps:
    levels = 3
    hop_fraction = [0.125, 0.5, 0.5]


def test_alignment_stitching():
    prior = StubPrior()
    total = 16
    z_top = torch.arange(2 * total).view(2, total)
    labels = dict(info=[dict(full_tokens=list(range(6))), dict(full_tokens=list(range(9)))])
    out = align.get_alignment(None, [None, None, z_top], labels, prior, False, Hps())
    starts = list(get_starts(total, prior.n_ctx, 4))
    assert starts == [0, 4, 8] and prior.calls == [0, 0, 4, 4, 8, 8]
    assert [a.shape for a in out] == [(16, 6), (16, 9)]
    for item, a in enumerate(out):
        n_full = a.shape[1]
        want = np.zeros((total, n_full + 1))
        for start in reversed(starts):
            _, idx = prior.get_y(labels, start)
            for r in range(prior.n_ctx):
                for j, col in enumerate(idx[item]):
                    want[start + r, col] = float(z_top[item, start + r]) + 0.25 * j
        assert np.array_equal(a, want[:, :-1])


def test_short_piece_is_padded_and_trimmed():
    prior = StubPrior()
    z_top = torch.arange(1, 6).view(1, 5)
    labels = dict(info=[dict(full_tokens=list(range(4)))])
    out = align.get_alignment(None, [None, None, z_top], labels, prior, True, Hps())
    assert out[0].shape == (5, 4)
    z, pad = align.pad_to_context(z_top, 8)
    assert pad == 3 and z.shape == (1, 8) and int(z[0, 5:].abs().sum()) == 0
