"""SimplePrior: everything around the autoregressive model of one level - which upper-level codes and labels a
window is conditioned on, how lyric tokens enter (merged into the token sequence, or through a separate
encoder), and the hand-over to ConditionalAutoregressive2D (whose sampling loop runs on the decode engine).

The constructor signature, the sub-module names (they are state-dict keys: `prior`, `prime_prior`,
`prime_state_proj`, `prime_state_ln`, `prime_x_out`, `conditioner_blocks`, `y_emb`) and the attributes that
sample.py / train.py / the notebooks read (n_ctx, raw_to_tokens, n_tokens, labeller, level(s), z_shapes,
downsamples, cond_downsample, sample_length, prior_dims, ...) follow the reference (jukebox/prior/prior.py);
the methods are written around two small helpers:

  TokenSpaces  - the shared vocabulary of a `single_enc_dec` prior: lyric tokens and VQ codes live in one sequence,
                 VQ codes shifted past the lyric vocabulary (reference :168-203)
  LyricEncoder - the separate lyric encoder of an encoder-decoder prior: tokens -> activations -> projection ->
                 LayerNorm = the keys/values of the decoder's enc-dec attention layers (reference :285-301)

z_forward / forward evaluate the loss (bits per token), predictions and recorded attention weights of a full window
without gradients; optimisation itself is out of scope.
"""
import numpy as np
import torch as t
import torch.nn as nn
import torch.nn.functional as F

from ..utils import dist_adapter as dist
from ..utils.dist_adapter import print_once
from ..utils.torch_utils import assert_shape
from ..transformer.ops import LayerNorm, Conv1D
from ..transformer import f32
from ..data.labels import EmptyLabeller, Labeller
from ..vqvae.vqvae import calculate_strides
from .autoregressive import ConditionalAutoregressive2D
from .conditioners import Conditioner, LabelConditioner


class TokenSpaces:
    """Several token streams laid end to end in one sequence with disjoint id ranges."""

    def __init__(self, shapes, bins, width):
        self.shapes = [tuple(s) for s in shapes]
        self.bins = [int(b) for b in bins]
        self.dims = [int(np.prod(s)) for s in self.shapes]
        self.shift = [int(v) for v in np.cumsum([0, *self.bins])[:-1]]
        self.width = width

    def merge(self, streams, conds):
        """streams: LongTensors [N, ...], one per leading space; conds: per space [N, dims, width] or None (zeros).
        Returns (tokens [N, sum dims of the given streams], cond [N, sum dims, width])."""
        N = streams[0].shape[0]
        toks = []
        for i, x in enumerate(streams):
            assert x.dtype == t.long, x.dtype
            assert (0 <= x).all() and (x < self.bins[i]).all(), f"token stream {i} outside [0, {self.bins[i]})"
            toks.append(x.reshape(N, -1) + self.shift[i])
        parts = []
        for i, c in enumerate(conds):
            if c is None:
                c = t.zeros((N, self.dims[i], self.width), dtype=t.float, device=streams[0].device)
            else:
                assert_shape(c, (N, self.dims[i], self.width))
            parts.append(c)
        return t.cat(toks, dim=1), t.cat(parts, dim=1)

    def last(self, z):
        """the last space's tokens of a merged sequence (which may be shorter than full), ids un-shifted"""
        N = z.shape[0]
        lead = sum(self.dims[:-1])
        x = z[:, lead:] - self.shift[-1]
        x = t.clamp(x, min=0)           # a sampled id below the last space's range (a lyric id) maps to code 0
        assert (x < self.bins[-1]).all(), f"rank {dist.get_rank()}: id outside the {self.bins[-1]} codes"
        return x.reshape(N, -1, *self.shapes[-1][1:])


class SimplePrior(nn.Module):
    def __init__(self, z_shapes, l_bins, encoder, decoder, level, downs_t, strides_t, labels, prior_kwargs,
                 x_cond_kwargs, y_cond_kwargs, prime_kwargs, copy_input, labels_v3=False, merged_decoder=False,
                 single_enc_dec=False):
        super().__init__()
        self.use_tokens = prime_kwargs.pop('use_tokens')
        self.n_tokens = prime_kwargs.pop('n_tokens')
        self.prime_loss_fraction = prime_kwargs.pop('prime_loss_fraction')
        self.copy_input = copy_input
        if copy_input:
            prime_kwargs['bins'] = l_bins
        self.z_shapes, self.levels, self.level = z_shapes, len(z_shapes), level
        assert level < self.levels, f"Total levels {self.levels}, got level {level}"
        self.z_shape = z_shapes[level]
        self.l_bins = l_bins
        self.encoder, self.decoder = encoder, decoder       # bound methods of the VQ-VAE: not sub-modules
        self.cond_level = level + 1
        self.x_cond = level != self.levels - 1               # every level but the top sees the level above
        self.y_cond = labels
        self.single_enc_dec = single_enc_dec
        self._build_conditioning(z_shapes, l_bins, downs_t, strides_t, x_cond_kwargs, y_cond_kwargs)
        if single_enc_dec:
            self._build_joint(prime_kwargs, prior_kwargs)
        else:
            self._build_separate(prime_kwargs, prior_kwargs, merged_decoder)
        self.n_ctx = self.gen_loss_dims
        self.total_loss_dims = self.prime_loss_dims + self.gen_loss_dims
        self.downsamples = calculate_strides(strides_t, downs_t)
        self.cond_downsample = None if not self.x_cond else self.downsamples[level + 1]
        self.raw_to_tokens = int(np.prod(self.downsamples[:level + 1]))
        self.sample_length = self.n_ctx * self.raw_to_tokens
        self.labels_v3 = labels_v3 if labels else False
        self.labeller = Labeller(self.y_emb.max_bow_genre_size, self.n_tokens, self.sample_length,
                                 v3=labels_v3) if labels else EmptyLabeller()
        print(f"Level:{level}, Cond downsample:{self.cond_downsample}, Raw to tokens:{self.raw_to_tokens}, "
              f"Sample length:{self.sample_length}")

    # ---- construction ---------------------------------------------------------------------------------------
    def _build_conditioning(self, z_shapes, l_bins, downs_t, strides_t, x_cond_kwargs, y_cond_kwargs):
        if self.x_cond:
            print_once("Conditioning on 1 above level(s)")
            up = self.cond_level
            self.conditioner_blocks = nn.ModuleList([
                Conditioner(input_shape=z_shapes[up], bins=l_bins, down_t=downs_t[up], stride_t=strides_t[up],
                            **x_cond_kwargs)])
        if self.y_cond:
            self.n_time = self.z_shape[0]
            # an upsampler takes its timing from the codes above; the top level needs the time signal
            self.y_emb = LabelConditioner(n_time=self.n_time, include_time_signal=not self.x_cond, **y_cond_kwargs)

    def _build_joint(self, prime_kwargs, prior_kwargs):
        """lyrics and codes in ONE autoregressive sequence (1b_lyrics)"""
        spaces = TokenSpaces([(self.n_tokens,), prior_kwargs.pop('input_shape')],
                             [prime_kwargs['bins'], prior_kwargs.pop('bins')], prior_kwargs['width'])
        self.spaces = spaces
        self.prior_shapes, self.prior_bins, self.prior_dims = spaces.shapes, spaces.bins, spaces.dims
        self.prior_bins_shift, self.prior_width = np.asarray(spaces.shift), spaces.width
        print_once(f'Creating cond. autoregress with prior bins {spaces.bins}, dims {spaces.dims}, '
                   f'shift {self.prior_bins_shift}, input shape {sum(spaces.dims)}, input bins {sum(spaces.bins)}')
        self.prime_loss_dims, self.gen_loss_dims = spaces.dims
        self.prior = ConditionalAutoregressive2D(input_shape=(sum(spaces.dims),), bins=sum(spaces.bins),
                                                 x_cond=(self.x_cond or self.y_cond), y_cond=True,
                                                 prime_len=self.prime_loss_dims, **prior_kwargs)

    def _build_separate(self, prime_kwargs, prior_kwargs, merged_decoder):
        """codes only in the decoder; lyrics (if any) through their own encoder (5b_lyrics) or not at all"""
        self.prime_loss_dims = 0
        if self.n_tokens != 0 and self.use_tokens:
            self.prime_loss_dims = int(self.n_tokens)
            self.prime_acts_width, self.prime_state_width = prime_kwargs['width'], prior_kwargs['width']
            self.prime_prior = ConditionalAutoregressive2D(input_shape=(self.n_tokens,), x_cond=False, y_cond=False,
                                                           only_encode=True, **prime_kwargs)
            self.prime_state_proj = Conv1D(self.prime_acts_width, self.prime_state_width,
                                           init_scale=prime_kwargs['init_scale'])
            self.prime_state_ln = LayerNorm(self.prime_state_width)
            self.prime_bins = prime_kwargs['bins']
            self.prime_x_out = nn.Linear(self.prime_state_width, self.prime_bins, bias=False)
            nn.init.normal_(self.prime_x_out.weight, std=0.02 * prior_kwargs['init_scale'])
        self.gen_loss_dims = int(np.prod(self.z_shape))
        self.prior = ConditionalAutoregressive2D(x_cond=(self.x_cond or self.y_cond), y_cond=self.y_cond,
                                                 encoder_dims=self.prime_loss_dims, merged_decoder=merged_decoder,
                                                 **prior_kwargs)

    @property
    def has_lyric_encoder(self):
        return (not self.single_enc_dec) and self.n_tokens != 0 and bool(self.use_tokens)

    # ---- what a window is conditioned on ----------------------------------------------------------------------
    def get_y(self, labels, start, get_indices=False):
        """label rows for the window whose first token is `start`: total length, offset of the window in raw samples,
        window length, artist, genres, and the lyric tokens that fall under the window"""
        if isinstance(self.labeller, EmptyLabeller):
            return None
        y = labels['y'].clone()
        y[:, 1] += int(start * self.raw_to_tokens)
        y[:, 2] = int(self.sample_length)
        indices = self.labeller.set_y_lyric_tokens(y, labels)
        return (y, indices) if get_indices else y

    def get_z_conds(self, zs, start, end):
        """codes of the level above under tokens [start, end) of this level (None at the top level)"""
        if not self.x_cond:
            return None
        ds = self.cond_downsample
        assert start % ds == 0 and end % ds == 0, f"window [{start},{end}) not aligned to {ds}"
        above = zs[self.level + 1][:, start // ds:end // ds]
        assert above.shape[1] == self.n_ctx // ds
        return [above]

    def x_emb(self, z_conds):
        """upper-level codes -> [N, n_ctx, width] through the conditioner stack (one block: one level above)"""
        blocks = self.conditioner_blocks
        z_conds = z_conds[:self.cond_level - self.level]
        assert len(z_conds) == len(blocks) == self.cond_level - self.level
        out = None
        for block, codes in zip(reversed(list(blocks)), reversed(list(z_conds))):
            out = block(codes, out)
        return out

    def get_cond(self, z_conds, y):
        """-> (x_cond [N, n_ctx, W] or [N, 1, W] or None, y_cond [N, 1, W] or None, lyric tokens or None)"""
        lyric = None
        if y is not None:
            n_labels = 4 + self.y_emb.max_bow_genre_size
            assert y.shape[1] == n_labels + self.n_tokens, \
                f"Expected {4} + {self.y_emb.max_bow_genre_size} + {self.n_tokens}, got {y.shape[1]}"
            y, lyric = y[:, :n_labels], y[:, n_labels:]
        y_cond = y_pos = None
        if self.y_cond:
            y_cond, y_pos = self.y_emb(y)
        x_cond = self.x_emb(z_conds) if self.x_cond else y_pos
        return x_cond, y_cond, lyric

    # single_enc_dec token-space helpers under the reference's names
    def prior_preprocess(self, xs, conds):
        return self.spaces.merge(xs, conds)

    def prior_postprocess(self, z):
        return self.spaces.last(z)

    # ---- VQ-VAE pass-through ----------------------------------------------------------------------------------
    def _level_span(self, start_level, end_level):
        return (self.level if start_level is None else start_level), (self.levels if end_level is None else end_level)

    def encode(self, x, start_level=None, end_level=None, bs_chunks=1):
        lo, hi = self._level_span(start_level, end_level)
        with t.no_grad():
            return self.encoder(x, start_level=lo, end_level=hi, bs_chunks=bs_chunks)

    def decode(self, zs, start_level=None, end_level=None, bs_chunks=1):
        lo, hi = self._level_span(start_level, end_level)
        assert len(zs) == hi - lo
        with t.no_grad():
            return self.decoder(zs, start_level=lo, end_level=hi, bs_chunks=bs_chunks)

    # ---- sampling -----------------------------------------------------------------------------------------------
    def sample(self, n_samples, z=None, z_conds=None, y=None, fp16=False, temp=1.0, top_k=0, top_p=0.0,
               chunk_size=None, sample_tokens=None):
        """one window: z = codes of this level already in the window (None / empty: ancestral), z_conds = codes of the
        level above, y = label rows.  Returns the codes [N, sample_tokens or n_ctx]."""
        for name, v in (("z", z), ("y", y), *((f"z_conds[{i}]", c) for i, c in enumerate(z_conds or []))):
            assert v is None or v.shape[0] == n_samples, f"{name}: expected batch {n_samples}, got {tuple(v.shape)}"
        fresh = z is None or z.shape[1] == 0
        if dist.get_rank() == 0:
            print(f"{'Ancestral' if fresh else 'Primed'} sampling {n_samples} samples with temp={temp}, "
                  f"top_k={top_k}, top_p={top_p}")
        how = dict(fp16=fp16, temp=temp, top_k=top_k, top_p=top_p)
        with t.no_grad():
            x_cond, y_cond, lyric = self.get_cond(z_conds, y)
            if self.single_enc_dec:
                out = self._sample_joint(n_samples, None if fresh else z, lyric, x_cond, y_cond, chunk_size, sample_tokens, how)
            else:
                out = self._sample_separate(n_samples, None if fresh else z, lyric, x_cond, y_cond, chunk_size, sample_tokens, how)
        if sample_tokens is None:
            assert_shape(out, (n_samples, *self.z_shape))
        return out

    def _sample_joint(self, N, z, lyric, x_cond, y_cond, chunk_size, sample_tokens, how):
        # the lyric tokens are the head of the sequence: always a primed run of the joint model
        given = [lyric] if z is None else [lyric, z]
        seq, cond = self.spaces.merge(given, [None, x_cond])
        total = None if sample_tokens is None else sample_tokens + self.n_tokens
        seq = self.prior.primed_sample(N, seq, cond, y_cond, chunk_size=chunk_size, sample_tokens=total, **how)
        return self.spaces.last(seq)

    def _sample_separate(self, N, z, lyric, x_cond, y_cond, chunk_size, sample_tokens, how):
        enc = self.get_encoder_kv(lyric, fp16=how["fp16"], sample=True)
        if z is None:
            return self.prior.sample(N, x_cond, y_cond, enc, sample_tokens=sample_tokens, **how)
        return self.prior.primed_sample(N, z, x_cond, y_cond, enc, chunk_size=chunk_size, sample_tokens=sample_tokens, **how)

    def get_encoder_kv(self, prime, fp16=False, sample=False):
        """lyric tokens -> encoder activations -> prime_state_proj (fp32 Conv1D) -> prime_state_ln: what the decoder's
        encoder-decoder attention layers read.  The reference parks the encoder on the CPU between windows to fit
        16 GB; with 180 GB it simply stays resident."""
        if not self.has_lyric_encoder:
            return None
        N = prime.shape[0]
        acts = self.prime_prior(prime, None, None, None, fp16=fp16)
        assert_shape(acts, (N, self.prime_loss_dims, self.prime_acts_width))
        assert acts.dtype == t.float
        states = f32.linear_kn(acts.reshape(-1, self.prime_acts_width), self.prime_state_proj.w,
                               self.prime_state_proj.b).view(N, self.prime_loss_dims, -1)
        kv = self.prime_state_ln(states)
        return kv.half() if (sample and fp16) else kv

    def get_prime_loss(self, encoder_kv, prime_t):
        """bits per lyric token of the encoder's next-token head (reference prior.py:303-310)"""
        if not self.use_tokens:
            return t.tensor(0.0, device=prime_t.device)
        N, L, W = encoder_kv.shape
        logits = f32.linear_nk(encoder_kv.float().reshape(N * L, W), self.prime_x_out.weight)
        return F.cross_entropy(logits, prime_t.reshape(-1)) / float(np.log(2.))

    def z_forward(self, z, z_conds=[], y=None, fp16=False, get_preds=False, get_attn_weights=False):
        """Evaluation forward over one full window of codes (reference prior.py:312-349): returns (loss, metrics), or -
        with get_attn_weights (True or a set of layer indices) - the recorded attention weights of those layers, which
        is what lyric alignment reads (reference jukebox/align.py get_alignment).
        No gradients are kept: optimisation is out of scope, the loss is the evaluation metric (bits per token)."""
        assert isinstance(get_attn_weights, (bool, set))
        tr = self.prior.transformer
        if get_attn_weights:
            tr.set_record_attn(get_attn_weights)
        x_cond, y_cond, lyric = self.get_cond(z_conds, y)
        if self.copy_input:
            lyric = z[:, :self.n_tokens]
        if self.single_enc_dec:
            seq, x_cond = self.prior_preprocess([lyric, z], [None, x_cond])
            (prime_loss, gen_loss), preds = self.prior(seq, x_cond, y_cond, fp16=fp16, get_sep_loss=True,
                                                       get_preds=get_preds)
        else:
            enc = self.get_encoder_kv(lyric, fp16=fp16)
            prime_loss = self.get_prime_loss(enc, lyric) if enc is not None else t.tensor(0.0, device=z.device)
            gen_loss, preds = self.prior(z, x_cond, y_cond, enc, fp16=fp16, get_preds=get_preds)
        if get_attn_weights:
            ws = tr.ws
            tr.set_record_attn(False)
            return ws
        total = self.total_loss_dims
        loss = self.prime_loss_fraction * prime_loss * self.prime_loss_dims / total + gen_loss * self.gen_loss_dims / total
        metrics = dict(bpd=gen_loss.clone(), prime_loss=prime_loss.clone(), gen_loss=gen_loss.clone())
        if get_preds:
            metrics["preds"] = preds.clone()
        return loss, metrics

    def forward(self, x, y=None, fp16=False, decode=False, get_preds=False):
        """audio -> codes of every level -> z_forward at this level (reference prior.py:351-359)"""
        z, *z_conds = self.encode(x, bs_chunks=x.shape[0])
        loss, metrics = self.z_forward(z=z, z_conds=z_conds, y=y, fp16=fp16, get_preds=get_preds)
        x_out = self.decode([z, *z_conds]) if decode else None
        return x_out, loss, metrics
