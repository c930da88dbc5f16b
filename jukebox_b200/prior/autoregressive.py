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
import torch.nn.functional as F

from ..transformer.ops import filter_logits_scaled, sample_categorical
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
        # the key lives ON the engine object: a freshly built engine (drop_engine after .cuda() /
        # load_state_dict) always starts without embeddings, whatever address CPython gives it
        key = (self.x_emb.weight.data_ptr(), self.pos_emb.pos_emb.data_ptr(),
               None if x_out is None else x_out.data_ptr(), None if start is None else start.data_ptr())
        if getattr(eng, "_emb_key", None) != key:
            eng.set_embeddings(x_emb=self.x_emb.weight, pos_emb=self.pos_emb.pos_emb, x_out=x_out,
                               start_token=None if start is None else start.view(-1))
            eng._emb_key = key
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
        cls = SamplingWindow if fp16 else SamplingWindowF32
        win = cls(self, n_samples, prime, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds,
                  sample_tokens)
        win.advance(win.sample_tokens)
        return win.finish()

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
        """Whole-sequence forward (reference autoregressive.py:116-172): the shifted token embeddings go through the
        transformer in forward mode; returns the activations for an `only_encode` model (the lyric encoder of
        separated enc-dec priors, prior.py:285-301), else (loss in bits per token, preds | acts | None).

        fp16=True runs the causal stack through the fp16 decode engine (prefill kernels); fp16=False runs the fp32
        forward-mode path (csrc/f32_path.cu).  No gradients: training is out of scope, the loss is an evaluation."""
        with t.no_grad():
            x = self.preprocess(x)
            N, D = x.shape
            assert D == self.input_dims, f"forward runs whole sequences of {self.input_dims} tokens, got {D}"
            assert (0 <= x).all() and (x < self.bins).all()
            x_cond, y_cond = self._check_conds(N, x_cond, y_cond)
            x = x.contiguous()
            if fp16 and not self.transformer._record_layers:
                acts = self._acts_fp16(x, x_cond, y_cond, encoder_kv)
            else:
                from ..transformer import f32
                h = f32.embed(self, x, y_cond, x_cond, N, D, 0)
                acts = self.transformer(h, encoder_kv=encoder_kv, fp16=False)
            if self.add_cond_after_transformer and x_cond is not None:
                acts = acts + x_cond
            if self.only_encode:
                return acts
            from ..transformer import f32
            preds = f32.linear_nk(acts.view(N * D, self.width), self.x_out.weight).view(N, D, self.bins)
            ln2 = float(np.log(2.))
            if get_sep_loss:
                assert self.prime_len is not None
                pl = self.prime_len
                loss = (F.cross_entropy(preds[:, :pl].reshape(-1, self.bins), x[:, :pl].reshape(-1)) / ln2,
                        F.cross_entropy(preds[:, pl:].reshape(-1, self.bins), x[:, pl:].reshape(-1)) / ln2)
            else:
                loss = F.cross_entropy(preds.view(-1, self.bins), x.view(-1)) / ln2
        if get_preds:
            return loss, preds
        if get_acts:
            return loss, acts
        return loss, None

    def _acts_fp16(self, x, x_cond, y_cond, encoder_kv):
        """the causal stack over given tokens on the fp16 decode engine: position by position it computes exactly the
        forward pass (reference check_sample, factored_attention.py:424-455)"""
        N, D = x.shape
        eng = self._engine(N)
        self.transformer.del_cache()
        if any(b.attn_func == 6 for b in self.transformer._attn_mods):
            assert encoder_kv is not None
            eng.set_encoder_kv(encoder_kv)
        acts = t.empty(N, D, self.width, dtype=t.float32, device=x.device)
        if 1 < D <= eng.prefill_capacity:
            eng.prefill(N, D, tokens=x, y_cond=y_cond, x_cond=x_cond, h_out=acts)
        else:
            for i in range(D):
                out = t.empty(N, self.width, dtype=t.float32, device=x.device)
                eng.step(N, tokens=x, y_cond=y_cond, x_cond=x_cond, h_out=out)
                acts[:, i] = out
        self.transformer.del_cache()
        return acts


class SamplingWindow:
    """One sampling window in flight on the decode engine: the body of the reference's sample loop
    (prior/autoregressive.py:222-237, 300-345) split into begin / advance / finish, so that callers which
    need the window in pieces (bench.py times 1/8-window slices) drive the same code as `sample`.

    begin (constructor): caches emptied, encoder K/V loaded, the given tokens prefilled in one pass when the
    engine can (else they are stepped by `advance`).  advance(upto): one decode launch + one sampling launch
    per position, nothing synchronises with the host.  finish(): cache bookkeeping + postprocess."""

    def __init__(self, ca, n_samples, prime, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds,
                 sample_tokens):
        assert ca.training is False
        assert not ca.only_encode
        assert fp16, "SamplingWindowF32 is the fp32 loop"
        self.ca = ca
        self.sample_tokens = ca.input_dims if sample_tokens is None else int(sample_tokens)
        self.N = N = n_samples
        self.x_cond, self.y_cond = ca._check_conds(N, x_cond, y_cond)
        self.P = P = prime.shape[1]
        assert P < self.sample_tokens <= ca.input_dims, \
            f"need given tokens {P} < sample_tokens {self.sample_tokens} <= input_dims {ca.input_dims}"
        dev = ca.x_emb.weight.device
        self.eng = eng = ca._engine(N)
        self.tr = tr = ca.transformer
        tr.del_cache()
        if any(b.attn_func == 6 for b in tr._attn_mods):
            assert encoder_kv is not None
            eng.set_encoder_kv(encoder_kv)
        self.tokens = t.zeros(N, self.sample_tokens, dtype=t.long, device=dev)
        if P:
            assert (0 <= prime).all() and (prime < ca.bins).all()
            self.tokens[:, :P] = prime
        self.get_preds = get_preds
        if get_preds:
            self.preds = t.empty(N, self.sample_tokens, ca.bins, dtype=t.float32, device=dev)
            self.lbuf, self.tstride = self.preds, ca.bins
        else:
            self.preds = None
            self.lbuf, self.tstride = t.empty(N, ca.bins, dtype=t.float32, device=dev), 0
        self.temp, self.top_k, self.top_p, self.fp16 = temp, top_k, top_p, fp16
        # the key of this call's Philox stream comes from torch's default generator, so t.manual_seed /
        # seed_per_rank make sampling reproducible exactly as they do for the reference's Categorical
        self.seed = int(t.empty((), dtype=t.int64).random_().item())
        self.pos = 0
        self.fbuf = None
        # x_cond is added BEHIND the stack (autoregressive.py:226-227) and the logits are linear in the activation:
        # x_cond . x_out^T of every position is computed once per window, so that the engine's logits product can take
        # the fp16-valued h on the tensor cores and add this bias in its epilogue (jkb200.h: jk_step_args.logit_bias)
        self.logit_bias = None
        if ca.add_cond_after_transformer and self.x_cond is not None and eng.has_logits_gemm:
            from ..transformer import f32
            with t.no_grad():
                Lc = self.x_cond.shape[1]
                self.logit_bias = f32.linear_nk(self.x_cond.reshape(N * Lc, ca.width), ca.x_out.weight).view(N, Lc, ca.bins)
        with t.no_grad():
            if 1 < P <= eng.prefill_capacity:
                # the given tokens go through all layers at once (the reference's chunked primed_sample,
                # autoregressive.py:300-338); chunk_size is moot - one chunk.  With get_preds their logits come from
                # the same pass: x_out over (activations + cond) in fp32 (autoregressive.py:318-325)
                if get_preds:
                    from ..transformer import f32
                    h = t.empty(N, P, ca.width, dtype=t.float32, device=dev)
                    eng.prefill(N, P, tokens=self.tokens, y_cond=self.y_cond, x_cond=self.x_cond, h_out=h)
                    if ca.add_cond_after_transformer and self.x_cond is not None:
                        h = h + (self.x_cond[:, :P] if self.x_cond.shape[1] > 1 else self.x_cond)
                    self.preds[:, :P] = f32.linear_nk(h.view(N * P, ca.width), ca.x_out.weight).view(N, P, ca.bins)
                else:
                    eng.prefill(N, P, tokens=self.tokens, y_cond=self.y_cond, x_cond=self.x_cond)
                self.pos = P

    def advance(self, upto):
        """positions [self.pos, upto): given positions are teacher-forced, the others sampled"""
        upto = min(int(upto), self.sample_tokens)
        eng, N, P, tokens = self.eng, self.N, self.P, self.tokens
        with t.no_grad():
            for sample_t in get_range(range(self.pos, upto)):
                need = self.get_preds or sample_t >= P
                eng.step(N, tokens=tokens, y_cond=self.y_cond, x_cond=self.x_cond,
                         logits=self.lbuf if need else None, logits_tstride=self.tstride, logit_bias=self.logit_bias)
                if sample_t >= P:
                    x = self.preds[:, sample_t] if self.get_preds else self.lbuf
                    if self.top_k or self.top_p:   # x / temp -> top-k / nucleus filter (ops.py:113-142): one launch
                        self.fbuf = filter_logits_scaled(x, self.temp, self.top_k, self.top_p, self.fbuf)
                        sample_categorical(self.fbuf, 1.0, self.seed, sample_t, tokens)
                    else:
                        sample_categorical(x, self.temp, self.seed, sample_t, tokens)
        self.pos = max(self.pos, upto)

    def finish(self):
        assert self.pos == self.sample_tokens, f"window stopped at {self.pos} of {self.sample_tokens}"
        tr = self.tr
        with t.no_grad():
            for b in tr._attn_mods:
                b.attn._advance(self.N, self.sample_tokens, self.fp16)
            tr.check_cache(self.N, self.sample_tokens, self.fp16)
            tr.del_cache()
            x = self.ca.postprocess(self.tokens, self.sample_tokens)
        return (x, self.preds) if self.get_preds else x


class SamplingWindowF32:
    """sample(fp16=False) / primed_sample(fp16=False): the same loop on the fp32 path (csrc/f32_path.cu) - embedding
    row, all layers on fp32 K/V caches, + cond, x_out in fp32, then the shared filter / Categorical kernels.  Four C-ABI
    calls per token instead of one persistent kernel: exactness path, not the hot path (train.py:139 sample logging)."""

    def __init__(self, ca, n_samples, prime, x_cond, y_cond, encoder_kv, fp16, temp, top_k, top_p, get_preds,
                 sample_tokens):
        assert ca.training is False and not ca.only_encode and not fp16
        self.ca = ca
        self.sample_tokens = ca.input_dims if sample_tokens is None else int(sample_tokens)
        self.N = N = n_samples
        self.x_cond, self.y_cond = ca._check_conds(N, x_cond, y_cond)
        self.P = P = prime.shape[1]
        assert P < self.sample_tokens <= ca.input_dims, \
            f"need given tokens {P} < sample_tokens {self.sample_tokens} <= input_dims {ca.input_dims}"
        dev = ca.x_emb.weight.device
        self.tr = ca.transformer
        self.tr.del_cache()
        self.encoder_kv = encoder_kv
        self.tokens = t.zeros(N, self.sample_tokens, dtype=t.long, device=dev)
        if P:
            assert (0 <= prime).all() and (prime < ca.bins).all()
            self.tokens[:, :P] = prime
        self.get_preds = get_preds
        self.preds = t.empty(N, self.sample_tokens, ca.bins, dtype=t.float32, device=dev) if get_preds else None
        self.temp, self.top_k, self.top_p = temp, top_k, top_p
        self.seed = int(t.empty((), dtype=t.int64).random_().item())
        self.pos = 0
        self.fbuf = None

    def advance(self, upto):
        from ..transformer import f32
        ca, N, P, tokens = self.ca, self.N, self.P, self.tokens
        upto = min(int(upto), self.sample_tokens)
        with t.no_grad():
            for sample_t in get_range(range(self.pos, upto)):
                self.tr.check_cache(N, sample_t, False)
                h = f32.embed(ca, tokens, self.y_cond, self.x_cond, N, 1, sample_t)
                h = self.tr(h, encoder_kv=self.encoder_kv, sample=True, fp16=False)
                if ca.add_cond_after_transformer and self.x_cond is not None:
                    h = h + (self.x_cond[:, sample_t:sample_t + 1] if self.x_cond.shape[1] > 1 else self.x_cond)
                if self.get_preds or sample_t >= P:
                    x = f32.linear_nk(h.view(N, ca.width), ca.x_out.weight)
                    if self.get_preds:
                        self.preds[:, sample_t] = x
                if sample_t >= P:
                    if self.top_k or self.top_p:
                        self.fbuf = filter_logits_scaled(x, self.temp, self.top_k, self.top_p, self.fbuf)
                        sample_categorical(self.fbuf, 1.0, self.seed, sample_t, tokens)
                    else:
                        sample_categorical(x, self.temp, self.seed, sample_t, tokens)
        self.pos = max(self.pos, upto)

    def finish(self):
        assert self.pos == self.sample_tokens, f"window stopped at {self.pos} of {self.sample_tokens}"
        with t.no_grad():
            self.tr.check_cache(self.N, self.sample_tokens, False)
            self.tr.del_cache()
            x = self.ca.postprocess(self.tokens, self.sample_tokens)
        return (x, self.preds) if self.get_preds else x
