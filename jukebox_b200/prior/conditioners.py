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
    def __init__(self, bins, out_width, init_scale):
        super().__init__()
        self.bins = bins
        self.emb = nn.Embedding(bins, out_width)
        nn.init.normal_(self.emb.weight, std=0.01 * init_scale)

    def forward(self, y):
        assert len(y.shape) == 2, f"Expected shape with 2 dims, got {y.shape}"
        assert y.dtype == t.long
        return _embed(y, self.emb.weight)


class RangeEmbedding(nn.Module):
    """positions in [pos_min, pos_max) binned into `bins` embeddings; with n_time > 1 the positions are
    interpolated between pos_start and pos_end (reference conditioners.py:70-111)."""

    def __init__(self, n_time, bins, range, out_width, init_scale, clamp=False):
        super().__init__()
        self.n_time, self.bins = n_time, bins
        self.emb = nn.Embedding(bins, out_width)
        nn.init.normal_(self.emb.weight, std=0.01 * init_scale)
        self.pos_min, self.pos_max = range
        self.clamp = clamp

    def forward(self, pos_start, pos_end=None):
        assert len(pos_start.shape) == 2
        pos_start = pos_start.float()
        if pos_end is not None:
            if self.clamp:
                pos_end = pos_end.clamp(self.pos_min, self.pos_max)
            pos_end = pos_end.float()
        if self.n_time != 1:
            assert pos_end is not None
            interp = t.arange(0, self.n_time, dtype=t.float, device=pos_start.device).view(1, self.n_time) / self.n_time
            position = pos_start + (pos_end - pos_start) * interp
        else:
            position = pos_start
        norm = (position - self.pos_min) / (self.pos_max - self.pos_min)
        bins = (self.bins * norm).floor().long().detach()
        return _embed(bins, self.emb.weight)


class LabelConditioner(nn.Module):
    def __init__(self, y_bins, t_bins, sr, min_duration, max_duration, n_time, out_width, init_scale,
                 max_bow_genre_size, include_time_signal):
        super().__init__()
        self.n_time, self.out_width = n_time, out_width
        assert len(y_bins) == 2, f"Expecting (genre, artist) bins, got {y_bins}"
        bow_genre_bins, artist_bins = y_bins
        self.max_bow_genre_size = max_bow_genre_size
        self.bow_genre_emb = SimpleEmbedding(bow_genre_bins, out_width, init_scale)
        self.artist_emb = SimpleEmbedding(artist_bins, out_width, init_scale)
        self.include_time_signal = include_time_signal
        if include_time_signal:
            self.total_length_emb = RangeEmbedding(1, t_bins, (min_duration * sr, max_duration * sr), out_width, init_scale)
            self.absolute_pos_emb = RangeEmbedding(n_time, t_bins, (0.0, max_duration * sr), out_width, init_scale)
            self.relative_pos_emb = RangeEmbedding(n_time, t_bins, (0.0, 1.0), out_width, init_scale, clamp=True)

    def forward(self, y):
        assert len(y.shape) == 2 and y.shape[-1] == 4 + self.max_bow_genre_size, f"bad label shape {y.shape}"
        assert y.dtype == t.long
        N = y.shape[0]
        total_length, offset, length, artist, genre = y[:, 0:1], y[:, 1:2], y[:, 2:3], y[:, 3:4], y[:, 4:]
        artist_emb = self.artist_emb(artist)
        mask = (genre >= 0).float().unsqueeze(2)          # empty genre slots are -1
        genre_emb = (self.bow_genre_emb(genre.clamp(0)) * mask).sum(dim=1, keepdim=True)
        start_emb = genre_emb + artist_emb
        assert_shape(start_emb, (N, 1, self.out_width))
        pos_emb = None
        if self.include_time_signal:
            start, end = offset, offset + length
            total_length, start, end = total_length.float(), start.float(), end.float()
            pos_emb = self.total_length_emb(total_length) + self.absolute_pos_emb(start, end) + \
                self.relative_pos_emb(start / total_length, end / total_length)
            assert_shape(pos_emb, (N, self.n_time, self.out_width))
        return start_emb, pos_emb
