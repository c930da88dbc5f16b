"""TEST INFRASTRUCTURE - plain-Python restatement of the KV-cache geometry, used only by tests/.

Two pieces of CUDA index arithmetic must agree with each other and with the reference's attention patterns
(factored_attention.py:135-228 sample branches, :328-353 _suff_cache_len):
  * decode: attn_geom() in jukebox_b200/csrc/decode_engine.cu - which cache rows position p attends and where its
    own K/V row is written;
  * chunked prefill: fwd_nkeys()/fwd_key() and kv_scatter_kernel in jukebox_b200/csrc/prefill.cu - which positions
    p attends inside the chunk and which cache row survives for it.
`reference_keys` states the reference's pattern directly in position space."""


def decode_geom(attn_func, p, bc, blocks, prime, enc_dims=0):
    """-> (first row, rows attended, current token included, row written or -1); decode_engine.cu attn_geom"""
    if attn_func == 0:
        return 0, p + 1, 1, p
    if attn_func == 1:
        return 0, p % bc + 1, 1, p % bc
    if attn_func == 2:
        base = (p % bc) * blocks
        return base, p // bc + 1, 1, base + p // bc
    if attn_func == 3:
        return ((p // bc + 1) & 1) * bc, (bc if p >= bc else 0), 0, ((p // bc) & 1) * bc + p % bc
    if attn_func == 7:
        return 0, min(p + 1, prime), (1 if p < prime else 0), (p if p < prime else -1)
    return 0, enc_dims, 0, -1


def prefill_keys(attn_func, p, bc, prime):
    """positions attended by p inside a prefill chunk; prefill.cu fwd_nkeys / fwd_key"""
    if attn_func == 0:
        return list(range(p + 1))
    if attn_func == 1:
        return [p - p % bc + j for j in range(p % bc + 1)]
    if attn_func == 2:
        return [p % bc + j * bc for j in range(p // bc + 1)]
    if attn_func == 3:
        return [(p // bc - 1) * bc + j for j in range(bc)] if p >= bc else []
    if attn_func == 7:
        return list(range(p + 1 if p < prime else prime))
    raise ValueError(attn_func)


def prefill_row(attn_func, p, P, bc, blocks, prime):
    """cache row that keeps position p after a prefill of P positions (-1: overwritten later / not cached);
    prefill.cu kv_scatter_kernel"""
    if attn_func == 0:
        return p
    if attn_func == 1:
        return p % bc if p + bc >= P else -1
    if attn_func == 2:
        return (p % bc) * blocks + p // bc
    if attn_func == 3:
        return ((p // bc) & 1) * bc + p % bc if p + 2 * bc >= P else -1
    if attn_func == 7:
        return p if p < prime else -1
    raise ValueError(attn_func)


def reference_keys(attn_func, p, bc, prime):
    """the reference's sample-mode pattern in position space (factored_attention.py:135-228)"""
    if attn_func == 0:                       # dense_attn: everything so far
        return list(range(p + 1))
    if attn_func == 1:                       # block_attn: own block, causal
        return list(range(p - p % bc, p + 1))
    if attn_func == 2:                       # transpose_block_attn: same offset in every block so far
        return list(range(p % bc, p + 1, bc))
    if attn_func == 3:                       # prev_block_attn: the whole previous block
        b = p // bc
        return list(range((b - 1) * bc, b * bc)) if b >= 1 else []
    if attn_func == 7:                       # prime_attn: the first _prime_len positions, causal inside them
        return list(range(min(p + 1, prime)))
    raise ValueError(attn_func)
