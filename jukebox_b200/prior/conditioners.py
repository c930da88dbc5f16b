"""Conditioning networks (reference: jukebox/prior/conditioners.py).

Conditioner: embed upper-level codes -> dilated-conv decoder block (x4 / x8 upsampling) ->
LayerNorm, all through libjkb200 kernels, channels-last.  LabelConditioner: artist / genre /
timing embeddings (a handful of table rows per sample; torch indexing on the GPU)."""
import torch as t
import torch.nn as nn

from ..transformer.ops import LayerNorm
from ..vqvae.encdec import DecoderConvBock
from ..utils.torch_utils import assert_shape
from .._lib import lib, check, ptr, stream_ptr


def _embed(idx, table, add=None):
    idx = idx.contiguous().long()
    flat = idx.view(-1)
    out = t.empty(flat.shape[0], table.shape[1], dtype=t.float32, device=idx.device)
    tab = table.detach().float().contiguous()
    check(lib().jk_embedding_f32(ptr(flat), ptr(tab), ptr(add), ptr(out), flat.shape[0], tab.shape[0], tab.shape[1],
                                 stream_ptr()))
    return out.view(*idx.shape, table.shape[1])


class Conditioner(nn.Module):
    def __init__(self, input_shape, bins, down_t, stride_t, out_width, init_scale, zero_out, res_scale, **block_kwargs):
        super().__init__()
        self.x_shape = input_shape
        self.width = out_width
        self.x_emb = nn.Embedding(bins, out_width)
        nn.init.normal_(self.x_emb.weight, std=0.02 * init_scale)
        self.cond = DecoderConvBock(self.width, self.width, down_t, stride_t, **block_kwargs, zero_out=zero_out,
                                    res_scale=res_scale)
        self.ln = LayerNorm(self.width)

    def forward(self, x, x_cond=None):
        N = x.shape[0]
        assert_shape(x, (N, *self.x_shape))
        add = None
        if x_cond is not None:
            assert_shape(x_cond, (N, *self.x_shape, self.width))
            add = x_cond.float().contiguous().view(-1, self.width)
        x = _embed(x, self.x_emb.weight, add)          # [N, T, width], already channels-last
        x = self.cond(x)
        return self.ln(x)


class SimpleEmbedding(nn.Module):
    """one table row per integer label (artist id, genre id)"""

    def __init__(self, bins, out_width, init_scale):
        super().__init__()
        self.bins = bins
        self.emb = nn.Embedding(bins, out_width)
        nn.init.normal_(self.emb.weight, std=0.01 * init_scale)

    def forward(self, y):
        assert y.dim() == 2 and y.dtype == t.long, f"expected a [N, k] LongTensor, got {tuple(y.shape)} {y.dtype}"
        lo, hi = int(y.min()), int(y.max())
        assert 0 <= lo and hi < self.bins, f"label ids must lie in [0, {self.bins}), got [{lo}, {hi}]"
        return _embed(y, self.emb.weight)


class RangeEmbedding(nn.Module):
    """A scalar position in [lo, hi) -> one of `bins` table rows (equal-width bins, right-open).  With n_time > 1
    the module embeds a whole window at once: n_time positions evenly spaced from pos_start towards pos_end
    (pos_end itself excluded), one row each - the timing signal of the top-level prior."""

    def __init__(self, n_time, bins, range, out_width, init_scale, clamp=False):
        super().__init__()
        self.n_time, self.bins = n_time, bins
        self.emb = nn.Embedding(bins, out_width)
        nn.init.normal_(self.emb.weight, std=0.01 * init_scale)
        self.pos_min, self.pos_max = range
        self.clamp = clamp

    def _check(self, pos, name, closed_right):
        assert pos.dim() == 2, f"{name}: expected [N, 1], got {tuple(pos.shape)}"
        lo, hi = float(pos.min()), float(pos.max())
        ok = self.pos_min <= lo and (hi <= self.pos_max if closed_right else hi < self.pos_max)
        assert ok, f"{name} outside [{self.pos_min}, {self.pos_max}{']' if closed_right else ')'}: [{lo}, {hi}]"

    def bin_ids(self, pos_start, pos_end=None):
        self._check(pos_start, "pos_start", closed_right=False)
        first = pos_start.float()
        if self.n_time == 1:
            where = first
        else:
            assert pos_end is not None, "a window needs its end position"
            if self.clamp:
                pos_end = pos_end.clamp(self.pos_min, self.pos_max)
            self._check(pos_end, "pos_end", closed_right=True)
            frac = t.arange(0, self.n_time, dtype=t.float, device=first.device).view(1, self.n_time) / self.n_time
            where = first + (pos_end.float() - first) * frac
        unit = (where - self.pos_min) / (self.pos_max - self.pos_min)          # [0, 1)
        return (self.bins * unit).floor().long()

    def forward(self, pos_start, pos_end=None):
        return _embed(self.bin_ids(pos_start, pos_end).detach(), self.emb.weight)


class LabelConditioner(nn.Module):
    """Label rows y = [total_length, offset, length, artist, genre_0 .. genre_{k-1}] (raw-sample units, genre slots
    padded with -1) -> (start embedding [N, 1, W] = artist + bag of genres, timing embedding [N, n_time, W] or None)."""

    COL_TOTAL, COL_OFFSET, COL_LENGTH, COL_ARTIST, COL_GENRE0 = 0, 1, 2, 3, 4

    def __init__(self, y_bins, t_bins, sr, min_duration, max_duration, n_time, out_width, init_scale,
                 max_bow_genre_size, include_time_signal):
        super().__init__()
        self.n_time, self.out_width = n_time, out_width
        assert len(y_bins) == 2, f"Expecting (genre, artist) bins, got {y_bins}"
        n_genres, n_artists = y_bins
        self.max_bow_genre_size = max_bow_genre_size
        self.bow_genre_emb = SimpleEmbedding(n_genres, out_width, init_scale)
        self.artist_emb = SimpleEmbedding(n_artists, out_width, init_scale)
        self.include_time_signal = include_time_signal
        if include_time_signal:
            longest = max_duration * sr
            self.total_length_emb = RangeEmbedding(1, t_bins, (min_duration * sr, longest), out_width, init_scale)
            self.absolute_pos_emb = RangeEmbedding(n_time, t_bins, (0.0, longest), out_width, init_scale)
            self.relative_pos_emb = RangeEmbedding(n_time, t_bins, (0.0, 1.0), out_width, init_scale, clamp=True)

    def start_embedding(self, y):
        artist = self.artist_emb(y[:, self.COL_ARTIST:self.COL_ARTIST + 1])
        genres = y[:, self.COL_GENRE0:]
        present = (genres >= 0).float().unsqueeze(2)            # empty genre slots are -1
        bag = (self.bow_genre_emb(genres.clamp(0)) * present).sum(dim=1, keepdim=True)
        return bag + artist

    def timing_embedding(self, y):
        total = y[:, self.COL_TOTAL:self.COL_TOTAL + 1].float()
        begin = y[:, self.COL_OFFSET:self.COL_OFFSET + 1]
        finish = (begin + y[:, self.COL_LENGTH:self.COL_LENGTH + 1]).float()
        begin = begin.float()
        return self.total_length_emb(total) + self.absolute_pos_emb(begin, finish) + \
            self.relative_pos_emb(begin / total, finish / total)

    def forward(self, y):
        assert y.dim() == 2 and y.shape[-1] == 4 + self.max_bow_genre_size, f"bad label shape {tuple(y.shape)}"
        assert y.dtype == t.long
        N = y.shape[0]
        start = self.start_embedding(y)
        assert_shape(start, (N, 1, self.out_width))
        timing = None
        if self.include_time_signal:
            timing = self.timing_embedding(y)
            assert_shape(timing, (N, self.n_time, self.out_width))
        return start, timing
