"""fp32 transformer path (csrc/f32_path.cu): Transformer.forward(fp16=False) in both modes.

  forward mode  - Transformer.forward(x, encoder_kv, sample=False): a whole sequence at once, the reference's training-shaped
                  call (transformer/transformer.py:169-192, masks of factored_attention.py:135-228), used by
                  ConditionalAutoregressive2D.forward (losses, get_preds, the only_encode lyric encoder) and by alignment
                  with record_attn (prior/prior.py:327-344);
  sampling mode - Transformer.forward(x, sample=True, fp16=False): positions appended one call at a time on fp32 K/V caches
                  (ConditionalAutoregressive2D.sample(fp16=False), which train.py:139 uses for its sample logging).

Not the hot path - the reference samples in fp16 (sample.py:239-241) and that is what the persistent decode kernel runs.
This path exists for exactness against the reference's fp32 outputs (tests at 2e-5) and as the fp32 side of the
reference's fp16-vs-fp32 criterion.  All arithmetic happens in libjkb200.so; there is no torch fallback.
"""
import ctypes as C

import torch as t

from .. import _lib


class F32Path:
    """Per-Transformer state of the fp32 path: fp32 parameter views, K/V caches of the current window, workspace."""

    def __init__(self, tr):
        self.tr = tr
        l0 = tr._attn_mods[0]
        self.dev = l0.ln_0.weight.device
        if self.dev.type != "cuda":
            raise RuntimeError("the fp32 transformer path needs the module on a CUDA device (jukebox_b200 has no CPU path)")
        self.W, self.S, self.M = tr.n_in, l0.attn.n_state, l0.mlp.c_fc.n_out
        self.depth = tr.n_depth
        self.keep = []          # fp32 copies of parameters stored in another dtype
        self.layers = (_lib.F32Layer * self.depth)()
        for i, blk in enumerate(tr._attn_mods):
            if blk.res_scale != 1.0:
                raise NotImplementedError("res_scale=True priors are not built (no named model uses them)")
            L = self.layers[i]
            a = blk.attn
            for name, p in (("ln0_g", blk.ln_0.weight), ("ln0_b", blk.ln_0.bias), ("ln1_g", blk.ln_1.weight),
                            ("ln1_b", blk.ln_1.bias), ("c_attn_w", a.c_attn.w), ("c_attn_b", a.c_attn.b),
                            ("c_proj_w", a.c_proj.w), ("c_proj_b", a.c_proj.b), ("fc_w", blk.mlp.c_fc.w),
                            ("fc_b", blk.mlp.c_fc.b), ("proj2_w", blk.mlp.c_proj.w), ("proj2_b", blk.mlp.c_proj.b)):
                setattr(L, name, self._f32(p))
            if blk.attn_func == 6:
                L.c_enc_kv_w, L.c_enc_kv_b = self._f32(a.c_enc_kv.w), self._f32(a.c_enc_kv.b)
            L.attn_func = blk.attn_func
        self.caches = None      # [(k, v)] per layer
        self.cache_n = 0
        self.pos = 0
        self.work = None

    def _f32(self, p):
        d = p.detach()
        if d.dtype != t.float32 or not d.is_contiguous():
            d = d.float().contiguous()
            self.keep.append(d)
        return d.data_ptr()

    def _alloc_caches(self, n):
        tr = self.tr
        self.caches = []
        for i, blk in enumerate(tr._attn_mods):
            rows = tr.encoder_dims if blk.attn_func == 6 else tr.n_ctx
            k = t.zeros(n, rows, self.S, dtype=t.float32, device=self.dev)
            v = t.zeros(n, rows, self.S, dtype=t.float32, device=self.dev)
            self.caches.append((k, v))
            self.layers[i].k_cache, self.layers[i].v_cache = k.data_ptr(), v.data_ptr()
        self.cache_n = n

    def reset(self):
        self.pos = 0
        self.caches = None
        self.cache_n = 0

    def run(self, x, encoder_kv, p0, record=None):
        """x: [n, P, width] fp32 CUDA (a new tensor is returned); positions [p0, p0 + P).  `record`: None or a list of
        layer indices whose attention weights are returned as {layer: [n, heads, P, keys]}."""
        tr = self.tr
        n, P, W = x.shape
        assert W == self.W
        if self.caches is None or self.cache_n != n:
            assert p0 == 0, "K/V caches of another batch size: call del_cache() first"
            self._alloc_caches(n)
        has6 = any(b.attn_func == 6 for b in tr._attn_mods)
        if has6 and p0 == 0:
            assert encoder_kv is not None and encoder_kv.shape == (n, tr.encoder_dims, W), \
                f"encoder_kv {None if encoder_kv is None else tuple(encoder_kv.shape)}, expected {(n, tr.encoder_dims, W)}"
            encoder_kv = encoder_kv.float().contiguous()
        else:
            encoder_kv = None
        out = x.float().contiguous().clone()
        ws = {}
        for i in range(self.depth):
            self.layers[i].attn_w = 0
        for i in (record or []):
            rows = tr.encoder_dims if tr._attn_mods[i].attn_func == 6 else tr.n_ctx
            ws[i] = t.empty(n, tr.n_head, P, rows, dtype=t.float32, device=self.dev)
            self.layers[i].attn_w = ws[i].data_ptr()
        a = _lib.F32Args(n=n, P=P, p0=p0, width=W, n_state=self.S, mlp_width=self.M, heads=tr.n_head, n_ctx=tr.n_ctx,
                         blocks=tr.blocks or 0, prime_len=tr.prime_len or 0, encoder_dims=tr.encoder_dims or 0,
                         depth=self.depth, x=out.data_ptr(), encoder_kv=_lib.ptr(encoder_kv).value or 0, work=0)
        if not has6:
            a.encoder_dims = 0
        need = C.c_size_t(0)
        _lib.check(_lib.lib().jk_f32_workspace_floats(C.byref(a), C.byref(need)))
        if self.work is None or self.work.numel() < need.value:
            self.work = t.empty(need.value, dtype=t.float32, device=self.dev)
        a.work = self.work.data_ptr()
        _lib.check(_lib.lib().jk_f32_forward(C.byref(a), self.layers, _lib.stream_ptr()))
        return out, ws


def embed(ca, tokens, y_cond, x_cond, n, P, p0):
    """[n, P, width] fp32 input rows of positions [p0, p0 + P) (csrc/f32_path.cu jk_f32_embed)."""
    dev = ca.x_emb.weight.device
    x = t.empty(n, P, ca.width, dtype=t.float32, device=dev)
    start = None if ca.y_cond else ca.start_token.detach().float().contiguous().view(-1)
    x_emb = ca.x_emb.weight.detach().float().contiguous()
    pos = ca.pos_emb.pos_emb.detach().float().contiguous()
    _lib.check(_lib.lib().jk_f32_embed(
        _lib.ptr(x), _lib.ptr(tokens), tokens.shape[1] if tokens is not None else 0, _lib.ptr(y_cond), _lib.ptr(x_cond),
        0 if x_cond is None else x_cond.shape[1], _lib.ptr(x_emb), _lib.ptr(pos), _lib.ptr(start), n, P, p0, ca.width,
        _lib.stream_ptr()))
    return x


def linear_nk(x, w):
    """x [M, K] . w[N, K]^T in fp32 (x_out, prior/autoregressive.py:86)"""
    M, K = x.shape
    N = w.shape[0]
    w = w.detach().float().contiguous()
    y = t.empty(M, N, dtype=t.float32, device=x.device)
    _lib.check(_lib.lib().jk_f32_linear(_lib.ptr(x.contiguous()), _lib.ptr(w), None, _lib.ptr(y), M, N, K, 1, _lib.stream_ptr()))
    return y


def linear_kn(x, w, b=None):
    """x [M, K] . w[K, N] + b in fp32 (Conv1D, transformer/ops.py:83-101)"""
    M, K = x.shape
    N = w.shape[1]
    w = w.detach().float().contiguous()
    b = None if b is None else b.detach().float().contiguous()
    y = t.empty(M, N, dtype=t.float32, device=x.device)
    _lib.check(_lib.lib().jk_f32_linear(_lib.ptr(x.float().contiguous()), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, N, K, 0,
                                        _lib.stream_ptr()))
    return y
