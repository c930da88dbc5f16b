"""Window planning / stitching of jukebox_b200.sample against the UNMODIFIED reference's own loop
(jukebox/sample.py:17-96), both driven with the same recording dummy prior on the CPU.  Skipped when the reference
tree is absent (GPU box)."""
import itertools

import pytest
import torch

from oracle.ref_import import reference_available, load_reference
from jukebox_b200.sample import plan_windows, Window
from jukebox_b200.utils.sample_utils import get_starts


class RecordingPrior:
    """prior.sample appends tokens that encode (call index, position), and records how it was called"""

    def __init__(self, n_ctx):
        self.n_ctx = n_ctx
        self.calls = []

    def get_z_conds(self, zs, start, end):
        return None

    def get_y(self, labels, start):
        return None

    def sample(self, n_samples, z=None, z_conds=None, y=None, sample_tokens=None, **kw):
        total = self.n_ctx if sample_tokens is None else sample_tokens
        self.calls.append((n_samples, z.shape[1], total, tuple(sorted(kw))))
        new = total - z.shape[1]
        assert new > 0
        fresh = 1000 * len(self.calls) + torch.arange(z.shape[1], total).view(1, -1).repeat(n_samples, 1)
        return torch.cat([z, fresh], dim=1)


class Hps(dict):
    __getattr__ = dict.__getitem__


CASES = [(total, n_ctx, hop, have, bs, mbs)
         for total, n_ctx, hop in [(40, 16, 8), (40, 16, 4), (16, 16, 8), (37, 16, 12), (10, 16, 8), (5, 16, 8), (33, 16, 16)]
         for have in (0, 3, 11, 16, 20) for bs, mbs in ((3, 2), (4, 4))]


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("total,n_ctx,hop,have,bs,mbs", CASES)
def test_sample_level_matches_reference(total, n_ctx, hop, have, bs, mbs):
    load_reference()
    import jukebox.sample as ref
    import jukebox_b200.sample as ours
    if total >= n_ctx and have > total:
        pytest.skip("more tokens than the level holds")
    outs = []
    for mod in (ref, ours):
        prior = RecordingPrior(n_ctx)
        zs = [torch.arange(have).view(1, -1).repeat(bs, 1)]
        hps = Hps(n_samples=bs)
        kw = dict(temp=0.9, fp16=True, max_batch_size=mbs)
        try:
            zs = mod.sample_level(zs, None, kw, 0, prior, total, hop, hps)
            outs.append((zs[0].clone(), prior.calls))
        except Exception as e:          # both sides must fail alike (e.g. negative slices)
            outs.append(("error", type(e).__name__))
    if outs[0][0] == "error" if isinstance(outs[0][0], str) else False:
        assert isinstance(outs[1][0], str)
        return
    assert not isinstance(outs[1][0], str), outs[1]
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]


def test_plan_windows_shapes():
    assert plan_windows(0, 40, 16, 8) == [Window(s, 16) for s in get_starts(40, 16, 8)]
    assert plan_windows(0, 10, 16, 8) == [Window(0, 10)]
    assert plan_windows(12, 10, 16, 8) == [Window(6, 16)]
    for total, n_ctx, hop in itertools.product((16, 17, 31, 64), (16,), (4, 8, 16)):
        wins = plan_windows(0, total, n_ctx, hop)
        assert wins[0].start == 0 and wins[-1].start + n_ctx == total
        assert all(b.start - a.start <= hop for a, b in zip(wins, wins[1:]))
