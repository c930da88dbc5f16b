"""SimplePrior: conditioning assembly around ConditionalAutoregressive2D.

Surface and attribute names follow the reference (jukebox/prior/prior.py) because sample.py,
train.py and the notebook drive the model through them: sample / encode / decode / get_z_conds /
get_y / get_cond / x_emb / prior_preprocess / prior_postprocess / get_encoder_kv, and n_ctx,
raw_to_tokens, n_tokens, labeller, level(s), z_shapes, downsamples, cond_downsample,
sample_length.  Training entry points (z_forward, forward, losses) are out of scope and raise.
"""
import numpy as np
import torch as t
import torch.nn as nn

from ..utils import dist_adapter as dist
from ..utils.dist_adapter import print_once
from ..utils.torch_utils import assert_shape
from ..transformer.ops import LayerNorm, Conv1D
from ..data.labels import EmptyLabeller, Labeller
from ..vqvae.vqvae import calculate_strides
from .autoregressive import ConditionalAutoregressive2D
from .conditioners import Conditioner, LabelConditioner


class SimplePrior(nn.Module):
    def __init__(self, z_shapes, l_bins, encoder, decoder, level, downs_t, strides_t, labels, prior_kwargs,
                 x_cond_kwargs, y_cond_kwargs, prime_kwargs, copy_input, labels_v3=False, merged_decoder=False,
                 single_enc_dec=False):
        super().__init__()
        self.use_tokens = prime_kwargs.pop('use_tokens')
        self.n_tokens = prime_kwargs.pop('n_tokens')
        self.prime_loss_fraction = prime_kwargs.pop('prime_loss_fraction')
        self.copy_input = copy_input
        if self.copy_input:
            prime_kwargs['bins'] = l_bins
        self.z_shapes = z_shapes
        self.levels = len(z_shapes)
        self.z_shape = z_shapes[level]
        self.level = level
        assert level < self.levels, f"Total levels {self.levels}, got level {level}"
        self.l_bins = l_bins
        # functions, not the vqvae module, so its parameters do not become ours
        self.encoder, self.decoder = encoder, decoder
        self.x_cond = (level != (self.levels - 1))
        self.cond_level = level + 1
        self.y_cond = labels
        self.single_enc_dec = single_enc_dec

        if self.x_cond:
            self.conditioner_blocks = nn.ModuleList()
            if dist.get_rank() == 0:
                print("Conditioning on 1 above level(s)")
            self.conditioner_blocks.append(Conditioner(input_shape=z_shapes[self.cond_level], bins=l_bins,
                                                       down_t=downs_t[self.cond_level],
                                                       stride_t=strides_t[self.cond_level], **x_cond_kwargs))
        if self.y_cond:
            self.n_time = self.z_shape[0]
            self.y_emb = LabelConditioner(n_time=self.n_time, include_time_signal=not self.x_cond, **y_cond_kwargs)

        if single_enc_dec:
            # lyric tokens and VQ codes share one sequence and one (shifted) vocabulary
            self.prior_shapes = [(self.n_tokens,), prior_kwargs.pop('input_shape')]
            self.prior_bins = [prime_kwargs['bins'], prior_kwargs.pop('bins')]
            self.prior_dims = [int(np.prod(shape)) for shape in self.prior_shapes]
            self.prior_bins_shift = np.cumsum([0, *self.prior_bins])[:-1]
            self.prior_width = prior_kwargs['width']
            print_once(f'Creating cond. autoregress with prior bins {self.prior_bins}, dims {self.prior_dims}, '
                       f'shift {self.prior_bins_shift}, input shape {sum(self.prior_dims)}, '
                       f'input bins {sum(self.prior_bins)}')
            self.prime_loss_dims, self.gen_loss_dims = self.prior_dims[0], self.prior_dims[1]
            self.total_loss_dims = self.prime_loss_dims + self.gen_loss_dims
            self.prior = ConditionalAutoregressive2D(input_shape=(sum(self.prior_dims),), bins=sum(self.prior_bins),
                                                     x_cond=(self.x_cond or self.y_cond), y_cond=True,
                                                     prime_len=self.prime_loss_dims, **prior_kwargs)
        else:
            if self.n_tokens != 0 and self.use_tokens:
                prime_input_shape = (self.n_tokens,)
                self.prime_loss_dims = int(np.prod(prime_input_shape))
                self.prime_acts_width, self.prime_state_width = prime_kwargs['width'], prior_kwargs['width']
                self.prime_prior = ConditionalAutoregressive2D(input_shape=prime_input_shape, x_cond=False,
                                                               y_cond=False, only_encode=True, **prime_kwargs)
                self.prime_state_proj = Conv1D(self.prime_acts_width, self.prime_state_width,
                                               init_scale=prime_kwargs['init_scale'])
                self.prime_state_ln = LayerNorm(self.prime_state_width)
                self.prime_bins = prime_kwargs['bins']
                self.prime_x_out = nn.Linear(self.prime_state_width, self.prime_bins, bias=False)
                nn.init.normal_(self.prime_x_out.weight, std=0.02 * prior_kwargs['init_scale'])
            else:
                self.prime_loss_dims = 0
            self.gen_loss_dims = int(np.prod(self.z_shape))
            self.total_loss_dims = self.prime_loss_dims + self.gen_loss_dims
            self.prior = ConditionalAutoregressive2D(x_cond=(self.x_cond or self.y_cond), y_cond=self.y_cond,
                                                     encoder_dims=self.prime_loss_dims, merged_decoder=merged_decoder,
                                                     **prior_kwargs)

        self.n_ctx = self.gen_loss_dims
        self.downsamples = calculate_strides(strides_t, downs_t)
        self.cond_downsample = self.downsamples[level + 1] if level != self.levels - 1 else None
        self.raw_to_tokens = int(np.prod(self.downsamples[:level + 1]))
        self.sample_length = self.n_ctx * self.raw_to_tokens
        if labels:
            self.labels_v3 = labels_v3
            self.labeller = Labeller(self.y_emb.max_bow_genre_size, self.n_tokens, self.sample_length, v3=self.labels_v3)
        else:
            self.labeller = EmptyLabeller()
        print(f"Level:{level}, Cond downsample:{self.cond_downsample}, Raw to tokens:{self.raw_to_tokens}, "
              f"Sample length:{self.sample_length}")

    # ---- per-window label / upper-level slices (reference :140-166) -----------------------------
    def get_y(self, labels, start, get_indices=False):
        if isinstance(self.labeller, EmptyLabeller):
            return None
        y = labels['y'].clone()
        y[:, 2] = int(self.sample_length)
        y[:, 1:2] = y[:, 1:2] + int(start * self.raw_to_tokens)
        indices = self.labeller.set_y_lyric_tokens(y, labels)
        return (y, indices) if get_indices else y

    def get_z_conds(self, zs, start, end):
        if self.level == self.levels - 1:
            return None
        assert start % self.cond_downsample == end % self.cond_downsample == 0
        z_cond = zs[self.level + 1][:, start // self.cond_downsample:end // self.cond_downsample]
        assert z_cond.shape[1] == self.n_ctx // self.cond_downsample
        return [z_cond]

    # ---- single_enc_dec token-space merge (reference :168-203) -----------------------------------
    def prior_preprocess(self, xs, conds):
        N = xs[0].shape[0]
        for i in range(len(xs)):
            bins, shift = int(self.prior_bins[i]), int(self.prior_bins_shift[i])
            assert xs[i].dtype == t.long, xs[i]
            assert (0 <= xs[i]).all() and (xs[i] < bins).all()
            xs[i] = (xs[i] + shift).view(N, -1)
        for i in range(len(conds)):
            dims = self.prior_dims[i]
            if conds[i] is not None:
                assert_shape(conds[i], (N, dims, self.prior_width))
            else:
                conds[i] = t.zeros((N, dims, self.prior_width), dtype=t.float, device=xs[0].device)
        return t.cat(xs, dim=1), t.cat(conds, dim=1)

    def prior_postprocess(self, z):
        N = z.shape[0]
        dims = (self.prior_dims[0], z.shape[1] - self.prior_dims[0])
        xs = list(t.split(z, dims, dim=1))
        for i in range(len(xs)):
            shape = self.prior_shapes[i]
            bins, shift = int(self.prior_bins[i]), int(self.prior_bins_shift[i])
            xs[i] = (xs[i] - shift).view(N, -1, *shape[1:])
            xs[i] = t.clamp(xs[i], min=0)   # sampled lyric tokens in the music range shift below 0
            assert (xs[i] < bins).all(), f'rank: {dist.get_rank()}, bins: {bins}, dims {dims}, shape {shape}'
        return xs[-1]

    def x_emb(self, z_conds):
        z_conds = z_conds[:self.cond_level - self.level]
        assert len(z_conds) == len(self.conditioner_blocks) == self.cond_level - self.level
        x_cond = None
        for z_cond, block in reversed(list(zip(z_conds, self.conditioner_blocks))):
            x_cond = block(z_cond, x_cond)
        return x_cond

    def encode(self, x, start_level=None, end_level=None, bs_chunks=1):
        start_level = self.level if start_level is None else start_level
        end_level = self.levels if end_level is None else end_level
        with t.no_grad():
            return self.encoder(x, start_level=start_level, end_level=end_level, bs_chunks=bs_chunks)

    def decode(self, zs, start_level=None, end_level=None, bs_chunks=1):
        start_level = self.level if start_level is None else start_level
        end_level = self.levels if end_level is None else end_level
        assert len(zs) == end_level - start_level
        with t.no_grad():
            return self.decoder(zs, start_level=start_level, end_level=end_level, bs_chunks=bs_chunks)

    def get_cond(self, z_conds, y):
        if y is not None:
            assert y.shape[1] == 4 + self.y_emb.max_bow_genre_size + self.n_tokens, \
                f"Expected {4} + {self.y_emb.max_bow_genre_size} + {self.n_tokens}, got {y.shape[1]}"
            n_labels = y.shape[1] - self.n_tokens
            y, prime = y[:, :n_labels], y[:, n_labels:]
        else:
            y, prime = None, None
        y_cond, y_pos = self.y_emb(y) if self.y_cond else (None, None)
        x_cond = self.x_emb(z_conds) if self.x_cond else y_pos
        return x_cond, y_cond, prime

    def sample(self, n_samples, z=None, z_conds=None, y=None, fp16=False, temp=1.0, top_k=0, top_p=0.0,
               chunk_size=None, sample_tokens=None):
        N = n_samples
        if z is not None:
            assert z.shape[0] == N, f"Expected shape ({N},**), got shape {z.shape}"
        if y is not None:
            assert y.shape[0] == N, f"Expected shape ({N},**), got shape {y.shape}"
        if z_conds is not None:
            for z_cond in z_conds:
                assert z_cond.shape[0] == N, f"Expected shape ({N},**), got shape {z_cond.shape}"
        no_past_context = (z is None or z.shape[1] == 0)
        if dist.get_rank() == 0:
            name = {True: 'Ancestral', False: 'Primed'}[no_past_context]
            print(f"{name} sampling {n_samples} samples with temp={temp}, top_k={top_k}, top_p={top_p}")
        with t.no_grad():
            x_cond, y_cond, prime = self.get_cond(z_conds, y)
            if self.single_enc_dec:
                if no_past_context:
                    z, x_cond = self.prior_preprocess([prime], [None, x_cond])
                else:
                    z, x_cond = self.prior_preprocess([prime, z], [None, x_cond])
                if sample_tokens is not None:
                    sample_tokens += self.n_tokens
                z = self.prior.primed_sample(n_samples, z, x_cond, y_cond, fp16=fp16, temp=temp, top_k=top_k,
                                             top_p=top_p, chunk_size=chunk_size, sample_tokens=sample_tokens)
                z = self.prior_postprocess(z)
            else:
                encoder_kv = self.get_encoder_kv(prime, fp16=fp16, sample=True)
                if no_past_context:
                    z = self.prior.sample(n_samples, x_cond, y_cond, encoder_kv, fp16=fp16, temp=temp, top_k=top_k,
                                          top_p=top_p, sample_tokens=sample_tokens)
                else:
                    z = self.prior.primed_sample(n_samples, z, x_cond, y_cond, encoder_kv, fp16=fp16, temp=temp,
                                                 top_k=top_k, top_p=top_p, chunk_size=chunk_size,
                                                 sample_tokens=sample_tokens)
            if sample_tokens is None:
                assert_shape(z, (N, *self.z_shape))
        return z

    def get_encoder_kv(self, prime, fp16=False, sample=False):
        """lyric encoder -> projection -> LayerNorm (reference :285-301).  The reference parks the
        encoder on the CPU between windows to fit 16 GB; with 180 GB it simply stays resident."""
        if self.n_tokens != 0 and self.use_tokens:
            N = prime.shape[0]
            prime_acts = self.prime_prior(prime, None, None, None, fp16=fp16)
            assert_shape(prime_acts, (N, self.prime_loss_dims, self.prime_acts_width))
            assert prime_acts.dtype == t.float
            # prime_state_proj is a Conv1D applied in fp32 (prior.py:294): a [N*L, w] x [w, W] product,
            # once per window
            w, b = self.prime_state_proj.w.float(), self.prime_state_proj.b.float()
            proj = t.addmm(b, prime_acts.view(-1, self.prime_acts_width), w).view(N, self.prime_loss_dims, -1)
            encoder_kv = self.prime_state_ln(proj)
            if sample and fp16:
                encoder_kv = encoder_kv.half()
            return encoder_kv
        return None

    def z_forward(self, *a, **k):
        raise NotImplementedError("training / alignment forward is out of scope (SURVEY.md section 2.1 #3)")

    def forward(self, *a, **k):
        raise NotImplementedError("training forward is out of scope (SURVEY.md section 2.1 #3)")
