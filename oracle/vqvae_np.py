"""CPU restatement (numpy, fp32) of the reference's VQ-VAE encode / quantise / decode path.

TEST INFRASTRUCTURE ONLY - "the oracle" (see oracle/transformer_np.py for the rules).
Pinned against the reference itself through tests/golden/vqvae_*.npz
(oracle/make_golden.py, tests/test_oracle_golden.py).

Restated (reference file:line under /root/reference/jukebox/vqvae/):
  conv1d / conv_transpose1d   torch.nn.Conv1d / ConvTranspose1d as used by encdec.py:17,41
  res_block                   resnet.py:27-44    x + res_scale * Conv1x1(ReLU(Conv3_dil(ReLU(x))))
  resnet1d                    resnet.py:46-75    dilation growth**(depth % cycle), optional reversal
  encoder_block/decoder_block encdec.py:6-46
  encoder/decoder             encdec.py:48-131   (decoder with all_levels=False)
  quantise                    bottleneck.py:112-119  argmin_j (|x|^2 - 2 x.k_j + |k_j|^2), fp32
  dequantise                  bottleneck.py:121-123, 138-147
  VQVAEOracle.encode/decode   vqvae.py:101-144

All tensors are NCT like the reference's internals; state-dict names are the reference's.
"""
import math
import numpy as np

F32 = np.float32


def conv1d(x, w, b, stride=1, pad=0, dil=1):
    """x:[N,C,T]  w:[O,C,K]  b:[O]  (torch.nn.Conv1d semantics)."""
    x = np.asarray(x, F32)
    N, C, T = x.shape
    O, _, K = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
    T_out = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
    y = np.zeros((N, O, T_out), F32)
    for k in range(K):
        xs = xp[:, :, k * dil: k * dil + (T_out - 1) * stride + 1: stride]
        y += np.einsum("oc,nct->not", w[:, :, k].astype(F32), xs, optimize=True)
    return y + b.astype(F32)[None, :, None]


def conv_transpose1d(x, w, b, stride=2, pad=1):
    """x:[N,C,T]  w:[C,O,K]  (torch.nn.ConvTranspose1d semantics, output_padding 0)."""
    x = np.asarray(x, F32)
    N, C, T = x.shape
    _, O, K = w.shape
    T_full = (T - 1) * stride + K
    y = np.zeros((N, O, T_full), F32)
    for k in range(K):
        y[:, :, k: k + (T - 1) * stride + 1: stride] += np.einsum(
            "co,nct->not", w[:, :, k].astype(F32), x, optimize=True)
    y = y[:, :, pad: T_full - pad]
    return y + b.astype(F32)[None, :, None]


def relu(x):
    return np.maximum(x, 0)


def resnet_dilations(depth, growth, cycle, reverse):
    d = [growth ** (i if cycle is None else i % cycle) for i in range(depth)]
    return d[::-1] if reverse else d


class _SD:
    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def __call__(self, name):
        return self.sd[self.prefix + name]

    def sub(self, p):
        return _SD(self.sd, self.prefix + p)


def resnet1d(x, P, depth, growth, cycle, reverse, res_scale, names="model"):
    """P: prefix accessor of the Resnet1D module; blocks live under `model.{j}` (Sequential)
    or `blocks.{j}` (checkpoint_res == 1, resnet.py:61-67)."""
    rs = 1.0 if not res_scale else 1.0 / math.sqrt(depth)
    for j, dil in enumerate(resnet_dilations(depth, growth, cycle, reverse)):
        B = P.sub(f"{names}.{j}.model.")
        h = conv1d(relu(x), B("1.weight"), B("1.bias"), 1, dil, dil)
        h = conv1d(relu(h), B("3.weight"), B("3.bias"), 1, 0, 1)
        x = x + F32(rs) * h
    return x


def decoder_block(x, P, down_t, stride_t, depth, growth, cycle, reverse, res_scale=False,
                  res_names="model"):
    """DecoderConvBock (encdec.py:28-46); P prefix ends with 'model.'"""
    if down_t == 0:
        return x
    x = conv1d(x, P("0.weight"), P("0.bias"), 1, 1, 1)
    for i in range(down_t):
        x = resnet1d(x, P.sub(f"{i + 1}.0."), depth, growth, cycle, reverse, res_scale, res_names)
        x = conv_transpose1d(x, P(f"{i + 1}.1.weight"), P(f"{i + 1}.1.bias"), stride_t, stride_t // 2)
    return x


def encoder_block(x, P, down_t, stride_t, depth, growth, cycle, res_scale=False):
    """EncoderConvBlock (encdec.py:6-26)."""
    if down_t == 0:
        return x
    for i in range(down_t):
        x = conv1d(x, P(f"{i}.0.weight"), P(f"{i}.0.bias"), stride_t, stride_t // 2, 1)
        x = resnet1d(x, P.sub(f"{i}.1."), depth, growth, cycle, False, res_scale)
    return conv1d(x, P(f"{down_t}.weight"), P(f"{down_t}.bias"), 1, 1, 1)


def quantise(x, k):
    """x:[M,w]  k:[bins,w]  -> int64 [M]  (bottleneck.py:112-119, expanded distance in fp32)."""
    x = np.asarray(x, F32)
    k = np.asarray(k, F32)
    dist = (x * x).sum(-1, keepdims=True, dtype=F32) - F32(2) * (x @ k.T) + (k * k).sum(-1, dtype=F32)[None]
    return dist.argmin(-1).astype(np.int64), dist


class VQVAEOracle:
    def __init__(self, sd, levels, downs_t, strides_t, width, depth, growth, cycle=None,
                 multipliers=None, reverse_decoder_dilation=True, emb_width=64):
        self.sd = {k: np.asarray(v) for k, v in sd.items()}
        self.levels, self.downs_t, self.strides_t = levels, downs_t, strides_t
        self.mult = multipliers or [1] * levels
        self.width, self.depth, self.growth, self.cycle = width, depth, growth, cycle
        self.reverse = reverse_decoder_dilation
        self.emb_width = emb_width

    def _encode_level(self, x, level):
        W, D = self.width * self.mult[level], self.depth * self.mult[level]
        for l in range(level + 1):
            P = _SD(self.sd, f"encoders.{level}.level_blocks.{l}.model.")
            x = encoder_block(x, P, self.downs_t[l], self.strides_t[l], D, self.growth, self.cycle)
        return x

    def encode_latents(self, x_ntc):
        """Pre-quantisation encoder outputs per level, each [N, emb, T_l]."""
        x = np.asarray(x_ntc, F32).transpose(0, 2, 1)
        return [self._encode_level(x, level) for level in range(self.levels)]

    def encode(self, x_ntc, start_level=0, end_level=None):
        end_level = self.levels if end_level is None else end_level
        zs = []
        for level, h in enumerate(self.encode_latents(x_ntc)):
            N, C, T = h.shape
            flat = h.transpose(0, 2, 1).reshape(-1, C)
            idx, _ = quantise(flat, self.sd[f"bottleneck.level_blocks.{level}.k"])
            zs.append(idx.reshape(N, T))
        return zs[start_level:end_level]

    def decode(self, zs, start_level=0, end_level=None):
        """zs[0] are the codes of `start_level`; only that level is used (vqvae.py:109-111)."""
        z = np.asarray(zs[0])
        k = self.sd[f"bottleneck.level_blocks.{start_level}.k"].astype(F32)
        x = k[z].transpose(0, 2, 1)                                   # dequantise -> NCT
        W, D = self.width * self.mult[start_level], self.depth * self.mult[start_level]
        for l in reversed(range(start_level + 1)):
            P = _SD(self.sd, f"decoders.{start_level}.level_blocks.{l}.model.")
            x = decoder_block(x, P, self.downs_t[l], self.strides_t[l], D, self.growth, self.cycle,
                              self.reverse)
        x = conv1d(x, self.sd[f"decoders.{start_level}.out.weight"],
                   self.sd[f"decoders.{start_level}.out.bias"], 1, 1, 1)
        return x.transpose(0, 2, 1)                                   # NTC


def conditioner(z, sd, prefix, down_t, stride_t, width, depth, growth, cycle, res_scale,
                x_cond=None, eps=1e-5):
    """Conditioner.forward (prior/conditioners.py:30-48): embed upper-level codes, run a
    DecoderConvBock (Resnet1D blocks registered as `blocks.{j}` because cond_c_res == 1),
    LayerNorm.  z:[N,T] int -> [N, T*stride**down, out_width] fp32."""
    from .transformer_np import layer_norm
    P = _SD(sd, prefix)
    x = P("x_emb.weight").astype(F32)[np.asarray(z)]
    if x_cond is not None:
        x = x + x_cond
    x = x.transpose(0, 2, 1)
    x = decoder_block(x, P.sub("cond.model."), down_t, stride_t, depth, growth, cycle, False,
                      res_scale, res_names="blocks")
    x = x.transpose(0, 2, 1)
    return layer_norm(x, P("ln.weight"), P("ln.bias"), eps)
