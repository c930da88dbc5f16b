"""VQVAE.encode / decode on the GPU (reference: jukebox/vqvae/vqvae.py:44-144).

Audio stays [N, T, 1] (the reference permutes to NCT for cuDNN; here every tensor is
channels-last, which is what the kernels want, so preprocess/postprocess are no-ops)."""
import numpy as np
import torch as t
import torch.nn as nn

from .encdec import Encoder, Decoder
from .bottleneck import Bottleneck


def calculate_strides(strides, downs):
    return [stride ** down for stride, down in zip(strides, downs)]


class VQVAE(nn.Module):
    def __init__(self, input_shape, levels, downs_t, strides_t, emb_width, l_bins, mu, commit, spectral,
                 multispectral, multipliers=None, use_bottleneck=True, **block_kwargs):
        super().__init__()
        assert use_bottleneck, "NoBottleneck variants are training-only experiments"
        self.sample_length = input_shape[0]
        x_shape, x_channels = input_shape[:-1], input_shape[-1]
        self.x_shape = x_shape
        self.downsamples = calculate_strides(strides_t, downs_t)
        self.hop_lengths = np.cumprod(self.downsamples)
        self.z_shapes = [(x_shape[0] // self.hop_lengths[level],) for level in range(levels)]
        self.levels = levels
        self.multipliers = [1] * levels if multipliers is None else multipliers
        assert len(self.multipliers) == levels, "Invalid number of multipliers"

        def kw(level):
            d = dict(block_kwargs)
            d["width"] *= self.multipliers[level]
            d["depth"] *= self.multipliers[level]
            return d
        self.encoders = nn.ModuleList(Encoder(x_channels, emb_width, level + 1, downs_t[:level + 1],
                                              strides_t[:level + 1], **kw(level)) for level in range(levels))
        self.decoders = nn.ModuleList(Decoder(x_channels, emb_width, level + 1, downs_t[:level + 1],
                                              strides_t[:level + 1], **kw(level)) for level in range(levels))
        self.bottleneck = Bottleneck(l_bins, emb_width, mu, levels)
        self.downs_t, self.strides_t, self.l_bins = downs_t, strides_t, l_bins
        self.commit, self.spectral, self.multispectral = commit, spectral, multispectral

    def preprocess(self, x):
        assert len(x.shape) == 3
        return x.float()

    def postprocess(self, x):
        return x

    def _decode(self, zs, start_level=0, end_level=None):
        if end_level is None:
            end_level = self.levels
        assert len(zs) == end_level - start_level
        xs_quantised = self.bottleneck.decode(zs, start_level=start_level, end_level=end_level)
        decoder, x_quantised = self.decoders[start_level], xs_quantised[0:1]
        return self.postprocess(decoder(x_quantised, all_levels=False))

    def decode(self, zs, start_level=0, end_level=None, bs_chunks=1):
        z_chunks = [t.chunk(z, bs_chunks, dim=0) for z in zs]
        x_outs = [self._decode([zc[i] for zc in z_chunks], start_level=start_level, end_level=end_level)
                  for i in range(bs_chunks)]
        return t.cat(x_outs, dim=0)

    def _encode(self, x, start_level=0, end_level=None):
        if end_level is None:
            end_level = self.levels
        x_in = self.preprocess(x)
        xs = [self.encoders[level](x_in)[-1] for level in range(self.levels)]
        return self.bottleneck.encode(xs)[start_level:end_level]

    def encode(self, x, start_level=0, end_level=None, bs_chunks=1):
        zs_list = [self._encode(x_i, start_level=start_level, end_level=end_level)
                   for x_i in t.chunk(x, bs_chunks, dim=0)]
        return [t.cat(z, dim=0) for z in zip(*zs_list)]

    def sample(self, n_samples):
        dev = self.bottleneck.level_blocks[0].k.device
        zs = [t.randint(0, self.l_bins, size=(n_samples, *z_shape), device=dev) for z_shape in self.z_shapes]
        return self.decode(zs)

    def forward(self, x, hps, loss_fn='l1'):
        raise NotImplementedError("VQ-VAE training (losses, codebook EMA) is out of scope")
