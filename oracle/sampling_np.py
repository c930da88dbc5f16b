"""TEST INFRASTRUCTURE - CPU restatement of the token-sampling step, used only by tests/.

Reference: jukebox/prior/autoregressive.py:233-235 (x / temp -> Categorical(logits=x).sample()).
The reference draws from torch's global generator; the product draws from Philox4x32-10 keyed per
call (csrc/sampling.cu), so parity here is (a) the generator itself against the published
Random123 known-answer vectors, (b) the inverse-CDF pick for a given uniform, and (c) the
distribution of many draws against softmax(logits / temp).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11).  ctr: 4 uint32, key: 2 uint32 -> 4 uint32."""
    c = [int(x) & MASK for x in ctr]
    k = [int(x) & MASK for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def uniform(seed, position, row):
    """the u in (2^-24, 1] behind (seed, position, row): counter (position, row, 'jk20', 0)."""
    r = philox4x32_10([position, row, 0x6A6B3230, 0], [seed & MASK, (seed >> 32) & MASK])[0]
    return np.float32(((r >> 8) + 1)) * np.float32(1.0 / 16777216.0)


def pick(logits, temp, u):
    """first bin whose CDF of softmax(logits / temp) reaches u (float64 accumulation)."""
    v = np.asarray(logits, np.float64) / temp
    e = np.where(np.isneginf(v), 0.0, np.exp(v - v[np.isfinite(v)].max()))
    cdf = np.cumsum(e)
    idx = int(np.searchsorted(cdf, float(u) * cdf[-1], side="left"))
    nz = np.nonzero(e)[0]
    return int(min(idx, nz[-1]))
