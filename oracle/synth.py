"""Deterministic synthetic weights, shared by oracle/make_golden.py, the tests and bench.py.

TEST / BENCH INFRASTRUCTURE.  The reference's default init (N(0, 0.02*init_scale), zero
biases, zero codebooks) makes every attention score and logit ~0, which would blind a
parity test to QK^T / softmax / argmin errors.  These weights keep
activations O(1), attention scores a few units wide and logits a few units wide.

A weight is a pure function of (name, shape, seed): numpy's legacy RandomState stream is
stable across numpy versions, so fixtures only need to store names, shapes and the seed.
"""
import zlib
import numpy as np


def _std_for(name, shape):
    last = name.split(".")[-1]
    if last == "w":                                   # Conv1D [n_in, n_out]
        gain = 1.5 if name.endswith("c_attn.w") else 1.0
        return gain / np.sqrt(shape[0])
    if last == "k":                                   # codebook
        return 1.0
    if "pos_emb" in name:
        return 0.5
    if last == "start_token":
        return 1.0
    if len(shape) == 3:                               # Conv1d [O,C,K] / ConvTranspose1d [C,O,K]
        return 1.0 / np.sqrt(shape[1] * shape[2])
    if len(shape) == 2:                               # embeddings / linear [out, in] or [bins, width]
        return 2.0 / np.sqrt(shape[1])
    return None


def synth_tensor(name, shape, seed=0):
    shape = tuple(int(s) for s in shape)
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    z = rs.standard_normal(shape).astype(np.float32)
    last = name.split(".")[-1]
    if len(shape) == 1:
        if last == "weight" and ("ln" in name.split(".")[-2] or name.split(".")[-2] == "ln"):
            return (1.0 + 0.1 * z).astype(np.float32)   # LayerNorm gamma
        return (0.1 * z).astype(np.float32)             # biases, LayerNorm beta
    return (z * np.float32(_std_for(name, shape))).astype(np.float32)


def synth_state_dict(named_shapes, seed=0):
    """named_shapes: iterable of (name, shape).  Returns {name: float32 ndarray}."""
    return {n: synth_tensor(n, s, seed) for n, s in named_shapes}
