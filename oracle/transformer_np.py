"""CPU restatement (numpy) of the reference's autoregressive-prior hot path.

TEST INFRASTRUCTURE ONLY - "the oracle".  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  The
product (jukebox_b200/) never does; it fails loudly without its CUDA library.

Parity pinning: the reference ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the reference itself,
run in the build container by oracle/make_golden.py and committed under
tests/golden/ (tests/test_oracle_golden.py checks every fixture).

What is restated (reference file:line, all under /root/reference/jukebox/):
  layer_norm          transformer/ops.py:14-24          (fp32 math, eps 1e-5)
  conv1d              transformer/ops.py:83-101         (y = x.w + b, w:[n_in,n_out])
  quick_gelu          transformer/ops.py:33-35          (x * sigmoid(1.702 x))
  attend_one          transformer/factored_attention.py:82-108 (q_len == 1: no mask)
  rows_attended       transformer/factored_attention.py:123-228, 328-353 (per-pattern
                      sample branches + _suff_cache_len, restated as the SET of
                      positions a query at position p attends)
  decode_layer        transformer/transformer.py:62-65,82-86 (ResAttnBlock sample branch)
  PriorOracle.step    prior/autoregressive.py:177-197,222-237 (get_emb, transformer,
                      +cond, x_out)
  forward_full        transformer/factored_attention.py:135-228 (forward-mode masks,
                      used for the encoder of separated enc-dec priors)

fp16 mode mirrors the reference's rounding points: activations
are fp16, every Conv1D output is rounded once to fp16 from an fp32 accumulator,
LayerNorm computes in fp32 and rounds its output, scores are rounded to fp16 before
AND after the 1/sqrt(dh) scaling, softmax is fp32, P is rounded to fp16, P.V is rounded
to fp16, the residual adds round to fp16.  Arrays are carried as float32 holding
fp16-representable values.
"""
import math
import numpy as np

F32 = np.float32


def r16(x):
    """Round to fp16 and widen back (value-preserving container is float32)."""
    return np.asarray(x, dtype=F32).astype(np.float16).astype(F32)


def maybe16(x, half):
    return r16(x) if half else np.asarray(x, dtype=F32)


def layer_norm(x, gamma, beta, eps=1e-5, half=False):
    # ops.py:20-24: input.float() -> layer_norm -> type_as(input)
    x = np.asarray(x, dtype=F32)
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    y = xc / np.sqrt(var + F32(eps)) * gamma.astype(F32) + beta.astype(F32)
    return maybe16(y, half)


def conv1d(x, w, b, half=False):
    # ops.py:99: addmm(b.type_as(x), x.view(-1, n_in), w.type_as(x))
    if half:
        w = r16(w)
        b = r16(b)
    y = np.matmul(np.asarray(x, dtype=F32), np.asarray(w, dtype=F32)) + np.asarray(b, dtype=F32)
    return maybe16(y, half)


def quick_gelu(x, half=False):
    # ops.py:33-35.  In fp16 the reference's eager CPU path rounds after each of the
    # three elementwise ops; restated exactly so.
    x = np.asarray(x, dtype=F32)
    if half:
        z = r16(F32(1.702) * x)
        s = r16(F32(1.0) / (F32(1.0) + np.exp(-z)))
        return r16(x * s)
    return x * (F32(1.0) / (F32(1.0) + np.exp(-F32(1.702) * x)))


def softmax32(w):
    w = np.asarray(w, dtype=F32)
    m = w.max(axis=-1, keepdims=True)
    e = np.exp(w - m)
    return e / e.sum(axis=-1, keepdims=True, dtype=F32)


def attend_one(q, K, V, n_head, half=False):
    """q:[bs,S]  K,V:[bs,L,S]  ->  [bs,S]   (factored_attention.py:82-121, q_len 1)."""
    bs, S = q.shape
    L = K.shape[1]
    dh = S // n_head
    qh = q.reshape(bs, n_head, 1, dh)
    Kh = K.reshape(bs, L, n_head, dh).transpose(0, 2, 3, 1)      # split_heads(k=True)
    Vh = V.reshape(bs, L, n_head, dh).transpose(0, 2, 1, 3)
    w = np.matmul(qh, Kh)                                         # fp16 GEMM: fp32 acc, one rounding
    w = maybe16(w, half)
    scale = 1.0 / math.sqrt(math.sqrt(dh))
    w = maybe16(w * F32(scale * scale), half)                     # w.mul_(scale*scale)
    p = maybe16(softmax32(w), half)                               # softmax in fp32, .type(wtype)
    a = maybe16(np.matmul(p, Vh), half)
    return a.transpose(0, 2, 1, 3).reshape(bs, S)                 # merge_heads


def prime_len_padded(prime_len, blocks):
    # factored_attention.py:303-308
    return (prime_len // blocks + 1) * blocks


def rows_attended(attn_func, p, block_ctx, prime_len_=None):
    """Positions (0-indexed) attended by the query at 0-indexed position p in decode.

    Returns ("rows", array) or ("zeros", n) for the prev-block first-block case
    (factored_attention.py:178-180: K = V = zeros(block_ctx))."""
    bc = block_ctx
    if attn_func == 0:
        return "rows", np.arange(0, p + 1)
    if attn_func == 1:
        return "rows", np.arange(p - p % bc, p + 1)
    if attn_func == 2:
        return "rows", np.arange(p % bc, p + 1, bc)
    if attn_func == 3:
        blk = p // bc
        if blk == 0:
            return "zeros", bc
        return "rows", np.arange((blk - 1) * bc, blk * bc)
    if attn_func == 7:
        return "rows", np.arange(0, min(p + 1, prime_len_))
    raise NotImplementedError(attn_func)


ATTN_ORDERS = {  # transformer/transformer.py:110-124
    0: lambda d: 0,
    1: lambda d: [1, 2][d % 2],
    2: lambda d: [1, 2, 3][d % 3],
    3: lambda d: [1, 4][d % 2],
    4: lambda d: [1, 5][d % 2],
    5: lambda d: [1, 4, 1, 1][d % 4],
    6: lambda d: [1, 2, 3, 6][d % 4],
    7: lambda d: [*[1, 2, 3] * 5, 6][d % 16],
    8: lambda d: [1, 2, 3, 1, 2, 3, 1, 2, 3, 6][d % 10],
    9: lambda d: [1, 2, 3, 0][d % 4],
    10: lambda d: [*[1, 2, 3, 1, 2, 3, 1, 2, 3], *[1, 2, 3, 1, 2, 3, 1, 2, 3, 6] * 7][d % 79],
    11: lambda d: [6, 6, 0][d % 3] if d % 16 == 15 else [1, 2, 3][d % 3],
    12: lambda d: [7, 7, 0][d % 3] if d % 16 == 15 else [1, 2, 3][d % 3],
}


class TransformerOracle:
    """Decode-mode (sample=True, one token at a time) transformer stack.

    `sd` maps reference state_dict names (relative to the Transformer module, i.e.
    `_attn_mods.{d}.attn.c_attn.w` ...) to numpy arrays."""

    def __init__(self, sd, n_in, n_ctx, n_head, n_depth, attn_order=0, blocks=None,
                 encoder_dims=None, prime_len=None, m_attn=0.25, res_scale=False):
        self.sd = {k: np.asarray(v) for k, v in sd.items()}
        self.n_in, self.n_ctx, self.n_head, self.n_depth = n_in, n_ctx, n_head, n_depth
        self.n_state = int(m_attn * n_in)
        self.blocks = blocks
        self.block_ctx = n_ctx // blocks if blocks else None
        self.encoder_dims = encoder_dims
        self.prime_len = prime_len
        self.attn_funcs = [ATTN_ORDERS[attn_order](d) for d in range(n_depth)]
        self.res_scale = (1.0 / n_depth) if res_scale else 1.0
        self.reset()

    def reset(self, max_len=None):
        self.max_len = self.n_ctx if max_len is None else min(self.n_ctx, max_len)
        self.K = [None] * self.n_depth
        self.V = [None] * self.n_depth
        self.enc_kv = [None] * self.n_depth
        self.t = 0

    def _p(self, d, name):
        return self.sd[f"_attn_mods.{d}.{name}"]

    def _attn(self, d, u, p, encoder_kv, half):
        af = self.attn_funcs[d]
        S = self.n_state
        bs = u.shape[0]
        if af == 6:
            # factored_attention.py:273-287
            q = conv1d(u, self._p(d, "attn.c_attn.w"), self._p(d, "attn.c_attn.b"), half)
            if self.enc_kv[d] is None:
                ekv = maybe16(encoder_kv, half)
                kv = conv1d(ekv, self._p(d, "attn.c_enc_kv.w"), self._p(d, "attn.c_enc_kv.b"), half)
                self.enc_kv[d] = (kv[..., :S], kv[..., S:])
            Ksel, Vsel = self.enc_kv[d]
        else:
            qkv = conv1d(u, self._p(d, "attn.c_attn.w"), self._p(d, "attn.c_attn.b"), half)
            q, k, v = qkv[:, :S], qkv[:, S:2 * S], qkv[:, 2 * S:]
            if self.K[d] is None:
                self.K[d] = np.zeros((bs, self.max_len, S), F32)
                self.V[d] = np.zeros((bs, self.max_len, S), F32)
            pl = prime_len_padded(self.prime_len, self.blocks) if af == 7 else None
            if af != 7 or p < pl:          # prime_qkv appends only while cache shorter than _prime_len
                self.K[d][:, p] = k
                self.V[d][:, p] = v
            kind, rows = rows_attended(af, p, self.block_ctx, pl)
            if kind == "zeros":
                Ksel = np.zeros((bs, rows, S), F32)
                Vsel = np.zeros((bs, rows, S), F32)
            else:
                Ksel, Vsel = self.K[d][:, rows], self.V[d][:, rows]
        a = attend_one(q, Ksel, Vsel, self.n_head, half)
        return conv1d(a, self._p(d, "attn.c_proj.w"), self._p(d, "attn.c_proj.b"), half)

    def step(self, x, encoder_kv=None, fp16=False):
        """x: [bs, n_in] fp32 -> [bs, n_in] fp32  (Transformer.forward, sample=True, 1 token)."""
        half = fp16
        h = maybe16(x, half)
        p = self.t
        for d in range(self.n_depth):
            u = layer_norm(h, self._p(d, "ln_0.weight"), self._p(d, "ln_0.bias"), half=half)
            a = self._attn(d, u, p, encoder_kv, half)
            x1 = maybe16(h + a, half)
            v = layer_norm(x1, self._p(d, "ln_1.weight"), self._p(d, "ln_1.bias"), half=half)
            g = quick_gelu(conv1d(v, self._p(d, "mlp.c_fc.w"), self._p(d, "mlp.c_fc.b"), half), half)
            m = conv1d(g, self._p(d, "mlp.c_proj.w"), self._p(d, "mlp.c_proj.b"), half)
            if self.res_scale == 1.0:
                h = maybe16(x1 + m, half)          # x + a + m, left to right
            else:
                h = maybe16(h + maybe16(F32(self.res_scale) * maybe16(a + m, half), half), half)
        self.t += 1
        return np.asarray(h, dtype=F32)

    # ---- forward (non-sample) mode: full sequence with masks; used for the lyric encoder ----
    def forward_full(self, x, encoder_kv=None, fp16=False):
        """x: [bs, L, n_in] with L == n_ctx.  Restates the forward-mode pattern math
        (factored_attention.py:135-228) by explicit masks; O(L^2) - small L only."""
        half = fp16
        h = maybe16(x, half)
        bs, L, _ = h.shape
        S, H = self.n_state, self.n_head
        dh = S // H
        pos = np.arange(L)
        for d in range(self.n_depth):
            af = self.attn_funcs[d]
            u = layer_norm(h, self._p(d, "ln_0.weight"), self._p(d, "ln_0.bias"), half=half)
            if af == 6:
                q = conv1d(u, self._p(d, "attn.c_attn.w"), self._p(d, "attn.c_attn.b"), half)
                kv = conv1d(maybe16(encoder_kv, half), self._p(d, "attn.c_enc_kv.w"),
                            self._p(d, "attn.c_enc_kv.b"), half)
                k, v = kv[..., :S], kv[..., S:]
                allow = np.ones((L, k.shape[1]), bool)
            else:
                qkv = conv1d(u, self._p(d, "attn.c_attn.w"), self._p(d, "attn.c_attn.b"), half)
                q, k, v = qkv[..., :S], qkv[..., S:2 * S], qkv[..., 2 * S:]
                bc = self.block_ctx
                qi, ki = pos[:, None], pos[None, :]
                if af == 0:
                    allow = ki <= qi
                elif af == 1:
                    allow = (ki <= qi) & (ki // bc == qi // bc)
                elif af == 2:
                    allow = (ki <= qi) & (ki % bc == qi % bc)
                elif af == 3:
                    allow = (ki // bc == qi // bc - 1)
                elif af == 7:
                    pl = prime_len_padded(self.prime_len, self.blocks)
                    allow = (ki <= qi) & (ki < pl)
                else:
                    raise NotImplementedError(af)
            qh = q.reshape(bs, L, H, dh).transpose(0, 2, 1, 3)
            kh = k.reshape(bs, -1, H, dh).transpose(0, 2, 3, 1)
            vh = v.reshape(bs, -1, H, dh).transpose(0, 2, 1, 3)
            w = maybe16(np.matmul(qh, kh), half)
            w = maybe16(w * F32(1.0 / math.sqrt(dh)), half)
            w = np.where(allow[None, None], w, F32(-1e9))
            if af == 3:
                # first block: every key of the zero block is visible -> uniform over bc zeros -> output 0
                first = (pos // self.block_ctx == 0)
                p_ = maybe16(softmax32(w), half)
                a = maybe16(np.matmul(p_, vh), half)
                a[:, :, first, :] = 0
            else:
                p_ = maybe16(softmax32(w), half)
                a = maybe16(np.matmul(p_, vh), half)
            a = a.transpose(0, 2, 1, 3).reshape(bs, L, S)
            a = conv1d(a, self._p(d, "attn.c_proj.w"), self._p(d, "attn.c_proj.b"), half)
            x1 = maybe16(h + a, half)
            vv = layer_norm(x1, self._p(d, "ln_1.weight"), self._p(d, "ln_1.bias"), half=half)
            g = quick_gelu(conv1d(vv, self._p(d, "mlp.c_fc.w"), self._p(d, "mlp.c_fc.b"), half), half)
            m = conv1d(g, self._p(d, "mlp.c_proj.w"), self._p(d, "mlp.c_proj.b"), half)
            h = maybe16(x1 + m, half)
        return np.asarray(h, dtype=F32)


class PriorOracle:
    """ConditionalAutoregressive2D decode loop with teacher-forced tokens -> logits.

    `sd` holds CA2D state_dict names: x_emb.weight, pos_emb.pos_emb, start_token (if not
    y_cond), x_out.weight (if untied), transformer._attn_mods....
    Restates prior/autoregressive.py:177-197 (get_emb) and :222-237 (sample loop body up to
    the logits; temperature/top-k/top-p/Categorical stay in torch on both sides)."""

    def __init__(self, sd, input_dims, bins, width, depth, heads, attn_order=0, blocks=None,
                 x_cond=False, y_cond=False, encoder_dims=0, merged_decoder=False,
                 prime_len=None, m_attn=0.25, res_scale=False, only_encode=False):
        self.sd = {k: np.asarray(v) for k, v in sd.items()}
        tsd = {k[len("transformer."):]: v for k, v in self.sd.items() if k.startswith("transformer.")}
        self.tr = TransformerOracle(tsd, width, input_dims, heads, depth, attn_order, blocks,
                                    encoder_dims, prime_len, m_attn, res_scale)
        self.input_dims, self.bins, self.width = input_dims, bins, width
        self.x_cond, self.y_cond = x_cond, y_cond
        self.add_cond_after = not merged_decoder
        self.only_encode = only_encode
        self.x_out = self.sd["x_emb.weight"] if not merged_decoder else self.sd.get("x_out.weight")

    def logits(self, tokens, x_cond=None, y_cond=None, encoder_kv=None, fp16=False, n_steps=None):
        """tokens: [bs, >= n_steps-1] int (token t-1 feeds step t).  Returns [bs, n_steps, bins]."""
        tokens = np.asarray(tokens)
        bs = tokens.shape[0]
        n_steps = n_steps or self.input_dims
        W = self.width
        if x_cond is None:
            x_cond = np.zeros((bs, 1, W), F32)
        self.tr.reset(max_len=n_steps)
        out = np.zeros((bs, n_steps, self.bins), F32)
        pos = self.sd["pos_emb.pos_emb"].astype(F32)
        emb = self.sd["x_emb.weight"].astype(F32)
        for t in range(n_steps):
            if t == 0:
                x = (np.asarray(y_cond, F32).reshape(bs, W) if self.y_cond
                     else np.broadcast_to(self.sd["start_token"].astype(F32), (bs, W)).copy())
            else:
                x = emb[tokens[:, t - 1]]
            cond = x_cond[:, t] if x_cond.shape[1] == self.input_dims else x_cond[:, 0]
            x = x + pos[t] + cond
            h = self.tr.step(x, encoder_kv=encoder_kv, fp16=fp16)
            if self.add_cond_after:
                h = h + cond
            out[:, t] = h @ self.x_out.astype(F32).T
        return out
