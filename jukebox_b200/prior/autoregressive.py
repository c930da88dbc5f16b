"""ConditionalAutoregressive2D: the sampling loop around the persistent decode kernel.

Reference surface kept (jukebox/prior/autoregressive.py): constructor signature, parameter names
(x_emb, pos_emb.pos_emb, start_token, transformer.*, x_out), sample(...), primed_sample(...),
preprocess / postprocess.  Per token the host enqueues ONE kernel (embedding gather + whole
transformer + fp32 logits, jukebox_b200/csrc/decode_engine.cu) and one sampling kernel
(temperature + Categorical, csrc/sampling.cu; top-k / top-p keep the reference's torch filter in
front of it) - nothing synchronises with the host inside the loop (the reference's per-token
`assert (0 <= x).all()` is hoisted out).
"""
import numpy as np
import torch as t
import torch.nn as nn

from ..transformer.ops import filter_logits, sample_categorical
from ..transformer.transformer import Transformer
from ..utils.logger import get_range


def get_normal(*shape, std=0.01):
    w = t.empty(shape)
    nn.init.normal_(w, std=std)
    return w


def split_chunks(length, chunk_size):
    n_passes = (length + chunk_size - 1) // chunk_size
    chunk_sizes = [*[chunk_size] * (n_passes - 1), (length - 1) % chunk_size + 1]
    assert sum(chunk_sizes) == length
    return chunk_sizes


class PositionEmbedding(nn.Module):
    def __init__(self, input_shape, width, init_scale=1.0, pos_init=False):
        super().__init__()
        assert not pos_init, "pos_init=True (factorised position embeddings) is not used by any named model"
        self.input_shape = input_shape
        self.input_dims = int(np.prod(input_shape))
        self.pos_init = pos_init
        self.pos_emb = nn.Parameter(get_normal(self.input_dims, width, std=0.01 * init_scale))

    def forward(self):
        return self.pos_emb


class ConditionalAutoregressive2D(nn.Module):
    def __init__(self, input_shape, bins, width=128, depth=2, heads=1, attn_dropout=0.0, resid_dropout=0.0,
                 emb_dropout=0.0, mask=True, zero_out=False, init_scale=1.0, res_scale=False, pos_init=False,
                 m_attn=0.25, m_mlp=1, checkpoint_res=0, checkpoint_attn=0, checkpoint_mlp=0, attn_order=0,
                 blocks=None, spread=None, x_cond=False, y_cond=False, encoder_dims=0, only_encode=False,
                 merged_decoder=False, prime_len=None):
        super().__init__()
        assert emb_dropout == 0.0, "dropout is a training feature"
        self.input_shape = input_shape
        self.input_dims = int(np.prod(input_shape))
        self.encoder_dims, self.bins, self.width, self.depth = encoder_dims, bins, width, depth
        self.x_emb = nn.Embedding(bins, width)
        nn.init.normal_(self.x_emb.weight, std=0.02 * init_scale)
        self.y_cond, self.x_cond = y_cond, x_cond
        if not y_cond:
            self.start_token = nn.Parameter(get_normal(1, width, std=0.01 * init_scale))
        self.pos_emb = PositionEmbedding(input_shape=input_shape, width=width, init_scale=init_scale, pos_init=pos_init)
        self.transformer = Transformer(n_in=width, n_ctx=self.input_dims, n_head=heads, n_depth=depth,
                                       attn_dropout=attn_dropout, resid_dropout=resid_dropout, afn='quick_gelu',
                                       scale=True, mask=mask, zero_out=zero_out, init_scale=init_scale,
                                       res_scale=res_scale, m_attn=m_attn, m_mlp=m_mlp,
                                       checkpoint_attn=checkpoint_attn, checkpoint_mlp=checkpoint_mlp,
                                       checkpoint_res=checkpoint_res, attn_order=attn_order, blocks=blocks,
                                       spread=spread, encoder_dims=encoder_dims, prime_len=prime_len)
        self.only_encode = only_encode
        self.prime_len = prime_len
        self.add_cond_after_transformer = not merged_decoder
        self.share_x_emb_x_out = not merged_decoder
        if not only_encode:
            self.x_out = nn.Linear(width, bins, bias=False)
            if self.share_x_emb_x_out:
                self.x_out.weight = self.x_emb.weight
        self._emb_key = None

    # ---- token <-> tensor layout (reference :100-112) -------------------------------------------
    def preprocess(self, x):
        return x.view(x.shape[0], -1).long()

    def postprocess(self, x, sample_tokens=None):
        N = x.shape[0]
        assert (0 <= x).all() and (x < self.bins).all()
        if sample_tokens is None or sample_tokens == self.input_dims:
            return x.view(N, *self.input_shape)
        return x.view(N, -1)

    # ---- engine plumbing ------------------------------------------------------------------
    def _engine(self, n_samples):
        tr = self.transformer
        tr.configure_engine(bins=0 if self.only_encode else self.bins,
                            add_cond_after=self.add_cond_after_transformer)
        eng = tr.engine(n_samples)
        x_out = None if self.only_encode else self.x_out.weight
        start = None if self.y_cond else self.start_token
        key = (id(eng), self.x_emb.weight.data_ptr(), self.pos_emb.pos_emb.data_ptr(),
               None if x_out is None else x_out.data_ptr(), None if start is None else start.data_ptr())
        if key != self._emb_key:
            eng.set_embeddings(x_emb=self.x_emb.weight, pos_emb=self.pos_emb.pos_emb, x_out=x_out,
                               start_token=None if start is None else start.view(-1))
            self._emb_key = key
        return eng

    def _check_conds(self, N, x_cond, y_cond):
        D = self.input_dims
        if self.y_cond:
            assert y_cond is not None
            assert y_cond.shape == (N, 1, self.width)
            y_cond = y_cond.float().contiguous().view(N, self.width)
        else:
            assert y_cond is None
        if self.x_cond:
            assert x_cond is not None
            assert x_cond.shape == (N, D, self.width) or x_cond.shape == (N, 1, self.width), \
                f"Got {x_cond.shape}, expected ({N}, {D}/{1}, {self.width})"
            x_cond = x_cond.float().contiguous()
        else:
            assert x_cond is None      # zeros in the reference; NULL for the kernel
        return x_cond, y_cond

    def _run(self, n_samples, prime, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds, sample_tokens):
        """shared body of sample / primed_sample.  prime: LongTensor [N, P] of given tokens (P may be 0)."""
        assert self.training is False
        assert not self.only_encode
        if not fp16:
            raise NotImplementedError("fp32 sampling is not built; use fp16=True (the reference's sampling_kwargs)")
        if sample_tokens is None:
            sample_tokens = self.input_dims
        N = n_samples
        x_cond, y_cond = self._check_conds(N, x_cond, y_cond)
        P = prime.shape[1]
        assert P < sample_tokens
        dev = self.x_emb.weight.device
        eng = self._engine(N)
        tr = self.transformer
        tr.del_cache()
        if any(b.attn_func == 6 for b in tr._attn_mods):
            assert encoder_kv is not None
            eng.set_encoder_kv(encoder_kv)
        tokens = t.zeros(N, sample_tokens, dtype=t.long, device=dev)
        if P:
            assert (0 <= prime).all() and (prime < self.bins).all()
            tokens[:, :P] = prime
        if get_preds:
            preds = t.empty(N, sample_tokens, self.bins, dtype=t.float32, device=dev)
            lbuf, tstride = preds, self.bins
        else:
            lbuf, tstride = t.empty(N, self.bins, dtype=t.float32, device=dev), 0
        # the key of this call's Philox stream comes from torch's default generator, so t.manual_seed /
        # seed_per_rank make sampling reproducible exactly as they do for the reference's Categorical
        seed = int(t.empty((), dtype=t.int64).random_().item())
        with t.no_grad():
            start = 0
            if P > 1 and not get_preds and 1 < P <= eng.prefill_capacity:
                # the given tokens go through all layers at once (the reference's chunked primed_sample,
                # autoregressive.py:300-338); chunk_size is moot - one chunk
                eng.prefill(N, P, tokens=tokens, y_cond=y_cond, x_cond=x_cond)
                start = P
            for sample_t in get_range(range(start, sample_tokens)):
                need = get_preds or sample_t >= P
                eng.step(N, tokens=tokens, y_cond=y_cond, x_cond=x_cond, logits=lbuf if need else None,
                         logits_tstride=tstride)
                if sample_t >= P:
                    x = preds[:, sample_t] if get_preds else lbuf
                    if top_k or top_p:      # filtering keeps the reference's torch expression (ops.py:99-122)
                        x = filter_logits(x / temp, top_k=top_k, top_p=top_p).contiguous()
                        sample_categorical(x, 1.0, seed, sample_t, tokens)
                    else:
                        sample_categorical(x, temp, seed, sample_t, tokens)
            for b in tr._attn_mods:
                b.attn._advance(N, sample_tokens, fp16)
            tr.check_cache(N, sample_tokens, fp16)
            tr.del_cache()
            x = self.postprocess(tokens, sample_tokens)
        return (x, preds) if get_preds else x

    # ---- reference API -----------------------------------------------------------------------
    def sample(self, n_samples, x_cond=None, y_cond=None, encoder_kv=None, fp16=False, temp=1.0, top_k=0,
               top_p=0.0, get_preds=False, sample_tokens=None):
        prime = t.zeros(n_samples, 0, dtype=t.long, device=self.x_emb.weight.device)
        return self._run(n_samples, prime, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds,
                         sample_tokens)

    def primed_sample(self, n_samples, x, x_cond=None, y_cond=None, encoder_kv=None, fp16=False, temp=1.0,
                      top_k=0, top_p=0.0, get_preds=False, chunk_size=None, sample_tokens=None):
        """`chunk_size` is accepted for compatibility: the prefill runs token by token through the
        same persistent kernel, which is what chunked prefill computes (reference check_chunks)."""
        with t.no_grad():
            x = self.preprocess(x)
        assert x.shape[0] == n_samples
        return self._run(n_samples, x, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds,
                         sample_tokens)

    def forward(self, x, x_cond=None, y_cond=None, encoder_kv=None, fp16=False, loss_full=False, encode=False,
                get_preds=False, get_acts=False, get_sep_loss=False):
        """Only the `only_encode` use is built (the lyric encoder of separated enc-dec priors,
        prior.py:285-301): returns activations [N, L, width].  The encoder is causal, so running it
        through the decode engine position by position computes exactly the full forward pass."""
        if not self.only_encode:
            raise NotImplementedError("training forward / losses are out of scope; forward-mode attention "
                                      "is the next hot-path row (SURVEY.md section 8f.1)")
        if not fp16:
            raise NotImplementedError("fp32 encoder path is not built; use fp16=True")
        with t.no_grad():
            x = self.preprocess(x)
            N, D = x.shape
            assert (0 <= x).all() and (x < self.bins).all()
            x_cond, y_cond = self._check_conds(N, x_cond, y_cond)
            eng = self._engine(N)
            self.transformer.del_cache()
            acts = t.empty(N, D, self.width, dtype=t.float32, device=x.device)
            x = x.contiguous()
            if 1 < D <= eng.prefill_capacity:
                eng.prefill(N, D, tokens=x, y_cond=y_cond, x_cond=x_cond, h_out=acts)
            else:
                for i in range(D):
                    out = t.empty(N, self.width, dtype=t.float32, device=x.device)
                    eng.step(N, tokens=x, y_cond=y_cond, x_cond=x_cond, h_out=out)
                    acts[:, i] = out
            self.transformer.del_cache()
            if self.add_cond_after_transformer and x_cond is not None:
                acts = acts + x_cond
        return acts
