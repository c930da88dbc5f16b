"""Torch restatement of the reference's decode step, for measuring fp16 ORDER NOISE on the GPU box.

TEST INFRASTRUCTURE ONLY (only tests/ may import it).  The reference cannot travel to the GPU box, so the question
"how far apart are two correct fp16 executions of this path?" is answered there by this module: the same torch
operators the reference calls, in the same order, at the same rounding points -
    LayerNorm      F.layer_norm(x.float(), ...).type_as(x)                transformer/ops.py:14-24
    Conv1D         torch.addmm(b.type_as(x), x, w.type_as(x))             transformer/ops.py:83-101
    attention      matmul -> mul_(scale*scale) -> softmax(float) -> type -> matmul   factored_attention.py:82-108
    rows attended  oracle.transformer_np.rows_attended (factored_attention.py:123-228, 328-353)
    block          x + a + m with quick_gelu (jit-scripted x * sigmoid(1.702 x))  transformer.py:19-30, 62-86
run on cuda in fp16 (cuBLAS HGEMM summation order) - i.e. what the reference itself computes on a GPU.  The GPU
tests compare |this - reference-CPU-fp16| (noise between two legitimate executions) with |ours - reference-CPU-fp16|.
It is pinned like the numpy oracle: tests/test_oracle_golden.py runs it on the CPU against the tiny fixtures.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .transformer_np import ATTN_ORDERS, rows_attended, prime_len_padded


class TorchDecodeOracle:
    def __init__(self, sd, n_in, n_ctx, n_head, n_depth, attn_order=0, blocks=None, encoder_dims=None, prime_len=None,
                 device="cpu", fp16_params=False):
        self.dev = torch.device(device)
        self.p = {}
        for k, v in sd.items():
            tsr = torch.from_numpy(np.asarray(v)).to(self.dev)
            if fp16_params and k.endswith(".w"):
                tsr = tsr.half()
            self.p[k] = tsr
        self.n_in, self.n_ctx, self.n_head, self.n_depth = n_in, n_ctx, n_head, n_depth
        self.n_state = n_in // 4
        self.blocks, self.block_ctx = blocks, (n_ctx // blocks if blocks else None)
        self.encoder_dims, self.prime_len = encoder_dims, prime_len
        self.attn_funcs = [ATTN_ORDERS[attn_order](d) for d in range(n_depth)]
        self.reset()

    def reset(self):
        self.K = [None] * self.n_depth
        self.V = [None] * self.n_depth
        self.enc = [None] * self.n_depth
        self.t = 0

    def _w(self, d, name):
        return self.p[f"_attn_mods.{d}.{name}"]

    @staticmethod
    def _conv1d(x, w, b):
        return torch.addmm(b.type_as(x), x, w.type_as(x))

    @staticmethod
    def _ln(x, g, b):
        return F.layer_norm(x.float(), (x.shape[-1],), g, b, 1e-5).type_as(x)

    def _attend(self, q, K, V):
        bs, S = q.shape
        H, dh = self.n_head, S // self.n_head
        qh = q.view(bs, 1, H, dh).permute(0, 2, 1, 3)
        kh = K.view(bs, -1, H, dh).permute(0, 2, 3, 1)
        vh = V.view(bs, -1, H, dh).permute(0, 2, 1, 3)
        scale = 1.0 / math.sqrt(math.sqrt(dh))
        w = torch.matmul(qh, kh)
        w.mul_(scale * scale)
        wtype = w.dtype
        w = F.softmax(w.float(), dim=-1).type(wtype)
        a = torch.matmul(w, vh)
        return a.permute(0, 2, 1, 3).contiguous().view(bs, S)

    def step(self, x, encoder_kv=None, fp16=True):
        """x: [bs, n_in] fp32 tensor on the oracle's device -> [bs, n_in] fp32"""
        h = x.half() if fp16 else x.float()
        p, S, bs = self.t, self.n_state, x.shape[0]
        for d in range(self.n_depth):
            af = self.attn_funcs[d]
            u = self._ln(h, self._w(d, "ln_0.weight"), self._w(d, "ln_0.bias"))
            qkv = self._conv1d(u, self._w(d, "attn.c_attn.w"), self._w(d, "attn.c_attn.b"))
            if af == 6:
                q = qkv
                if self.enc[d] is None:
                    e = encoder_kv.type_as(h).reshape(-1, self.n_in)
                    kv = self._conv1d(e, self._w(d, "attn.c_enc_kv.w"), self._w(d, "attn.c_enc_kv.b")).view(bs, -1, 2 * S)
                    self.enc[d] = (kv[..., :S].contiguous(), kv[..., S:].contiguous())
                Ks, Vs = self.enc[d]
            else:
                q, k, v = qkv[:, :S], qkv[:, S:2 * S], qkv[:, 2 * S:]
                if self.K[d] is None:
                    self.K[d] = torch.zeros(bs, self.n_ctx, S, dtype=h.dtype, device=self.dev)
                    self.V[d] = torch.zeros(bs, self.n_ctx, S, dtype=h.dtype, device=self.dev)
                pl = prime_len_padded(self.prime_len, self.blocks) if af == 7 else None
                if af != 7 or p < pl:
                    self.K[d][:, p] = k
                    self.V[d][:, p] = v
                kind, rows = rows_attended(af, p, self.block_ctx, pl)
                if kind == "zeros":
                    Ks = torch.zeros(bs, rows, S, dtype=h.dtype, device=self.dev)
                    Vs = torch.zeros_like(Ks)
                elif len(rows) and rows[-1] - rows[0] == len(rows) - 1:      # contiguous run: a view, like the reference's slices
                    Ks, Vs = self.K[d][:, rows[0]:rows[-1] + 1], self.V[d][:, rows[0]:rows[-1] + 1]
                else:
                    idx = torch.as_tensor(rows, device=self.dev)
                    Ks, Vs = self.K[d].index_select(1, idx), self.V[d].index_select(1, idx)
            a = self._attend(q.contiguous(), Ks, Vs)
            a = self._conv1d(a, self._w(d, "attn.c_proj.w"), self._w(d, "attn.c_proj.b"))
            v1 = self._ln(h + a, self._w(d, "ln_1.weight"), self._w(d, "ln_1.bias"))
            f = self._conv1d(v1, self._w(d, "mlp.c_fc.w"), self._w(d, "mlp.c_fc.b"))
            g = f * torch.sigmoid(1.702 * f)
            m = self._conv1d(g, self._w(d, "mlp.c_proj.w"), self._w(d, "mlp.c_proj.b"))
            h = h + a + m
        self.t += 1
        return h.float()
