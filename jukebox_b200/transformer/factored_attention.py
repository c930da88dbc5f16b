"""FactoredAttention: parameters + decode-time bookkeeping of one attention module.

Mirrors the public attributes of the reference class (jukebox/transformer/factored_attention.py:
attn_func, blocks, block_ctx, sample_t, cache, record_attn, w, _prime_len, _suff_cache_len,
check_cache, del_cache).  The KV cache itself lives inside the decode engine in a per-pattern
layout (DESIGN.md "KV cache"); `cache` therefore only reports the logical length.
"""
import torch as t
import torch.nn as nn

from .ops import Conv1D

ATTN_FUNCS = {0: "dense", 1: "block", 2: "transpose_block", 3: "prev_block", 4: "summary",
              5: "summary_spread", 6: "decode (enc-dec)", 7: "prime"}


class FactoredAttention(nn.Module):
    def __init__(self, n_in, n_ctx, n_state, n_head, attn_dropout=0.0, resid_dropout=0.0, scale=True,
                 mask=False, zero_out=False, init_scale=1.0, checkpoint_attn=0, attn_func=0, blocks=None,
                 spread=None, encoder_dims=None, prime_len=None):
        super().__init__()
        assert n_state % n_head == 0
        assert attn_dropout == 0.0 and resid_dropout == 0.0, "dropout is a training feature (out of scope)"
        self.n_in, self.n_ctx, self.n_state, self.n_head = n_in, n_ctx, n_state, n_head
        self.scale, self.mask = scale, mask
        if attn_func == 6:
            self.c_attn = Conv1D(n_in, n_state, init_scale=init_scale)
            self.c_enc_kv = Conv1D(n_in, n_state * 2, init_scale=init_scale)
        else:
            self.c_attn = Conv1D(n_in, n_state * 3, init_scale=init_scale)
        self.c_proj = Conv1D(n_state, n_in, zero_out, init_scale=init_scale)
        assert attn_func in ATTN_FUNCS
        self.attn_func = attn_func
        self.blocks, self.spread = blocks, spread
        if blocks is not None:
            assert n_ctx % blocks == 0
            self.block_ctx = n_ctx // blocks
        self.checkpoint_attn = checkpoint_attn
        self.sample_t = 0
        self.cache = {}
        self.encoder_dims = encoder_dims
        self.prime_len = prime_len
        self.record_attn = False
        self.w = None

    @property
    def _prime_len(self):
        assert self.prime_len is not None
        return (self.prime_len // self.blocks + 1) * self.blocks

    def _suff_cache_len(self):
        """rows of K/V a query at 1-indexed position sample_t needs (reference :328-353)."""
        s, f = self.sample_t, self.attn_func
        if f in (0, 2):
            return s
        if f == 1:
            return (s - 1) % self.block_ctx + 1
        if f == 3:
            return s if s <= self.block_ctx else (s - 1) % self.block_ctx + 1 + self.block_ctx
        if f == 6:
            return self.encoder_dims
        if f == 7:
            return min(s, self._prime_len)
        raise NotImplementedError(f"attn_func {f} ({ATTN_FUNCS[f]}) has no sampling path (same as the reference)")

    def check_cache(self, n_samples, sample_t, fp16):
        assert self.sample_t == sample_t, f"{self.sample_t} != {sample_t}"
        if sample_t == 0:
            assert self.cache == {}
        else:
            assert self.cache.get("n_samples") == n_samples
            assert self.cache.get("len") == self._suff_cache_len()
            assert self.cache.get("dtype") == (t.float16 if fp16 else t.float32)

    def del_cache(self):
        self.sample_t = 0
        self.cache = {}

    def _advance(self, n_samples, n_tokens, fp16):
        self.sample_t += n_tokens
        self.cache = dict(n_samples=n_samples, len=self._suff_cache_len(),
                          dtype=t.float16 if fp16 else t.float32)

    def forward(self, x, encoder_kv=None, sample=False):
        raise RuntimeError("FactoredAttention runs fused inside the decode engine; call "
                           "Transformer.forward(x, sample=True, fp16=True)")
