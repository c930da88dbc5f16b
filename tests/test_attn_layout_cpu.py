"""The decode kernel's cache geometry, the chunked prefill's key sets and its scatter rule agree with each other
and with the reference's patterns for every (pattern, block size, chunk length) - a sweep the GPU parity tests
only sample (oracle/attn_layout_np.py restates the CUDA index arithmetic line by line)."""
import itertools

import pytest

from oracle.attn_layout_np import decode_geom, prefill_keys, prefill_row, reference_keys


def _rows(attn_func, n_ctx, bc, prime):
    return {0: n_ctx, 1: bc, 2: n_ctx, 3: 2 * bc, 7: prime}[attn_func]


@pytest.mark.parametrize("attn_func", [0, 1, 2, 3, 7])
def test_decode_geometry_reads_the_reference_pattern(attn_func):
    for bc, blocks in ((4, 6), (12, 8), (7, 5)):
        n_ctx = bc * blocks
        prime = 2 * bc + 1 if attn_func == 7 else 0
        cache = [None] * _rows(attn_func, n_ctx, bc, prime)          # row -> position stored there
        for p in range(n_ctx):
            base, R, cur, wrow = decode_geom(attn_func, p, bc, blocks, prime)
            if cur and wrow >= 0:
                cache[wrow] = p                                         # the kernel appends the current row to its tile
            seen = [cache[base + i] for i in range(R - cur)] + ([p] if cur else [])
            assert sorted(seen) == reference_keys(attn_func, p, bc, prime), (attn_func, bc, p)
            if not cur and wrow >= 0:
                cache[wrow] = p                                         # patterns that do not attend p still cache it


@pytest.mark.parametrize("attn_func", [0, 1, 2, 3, 7])
def test_prefill_matches_stepping(attn_func):
    for (bc, blocks), in itertools.product(((4, 6), (12, 8), (7, 5))):
        n_ctx = bc * blocks
        prime = 2 * bc + 1 if attn_func == 7 else 0
        rows = _rows(attn_func, n_ctx, bc, prime)
        for P in range(1, n_ctx):
            # key sets inside the chunk
            for p in range(P):
                assert prefill_keys(attn_func, p, bc, prime) == reference_keys(attn_func, p, bc, prime)
            # cache state after P decode steps ...
            stepped = [None] * rows
            for p in range(P):
                wrow = decode_geom(attn_func, p, bc, blocks, prime)[3]
                if wrow >= 0:
                    stepped[wrow] = p
            # ... equals the scatter of the survivors (each row written by exactly one position)
            scattered = [None] * rows
            for p in range(P):
                r = prefill_row(attn_func, p, P, bc, blocks, prime)
                if r >= 0:
                    assert scattered[r] is None, "two positions scatter to one row"
                    scattered[r] = p
            assert scattered == stepped, (attn_func, bc, P)
            # and the next decode step reads what the reference would
            base, R, cur, _ = decode_geom(attn_func, P, bc, blocks, prime)
            seen = [scattered[base + i] for i in range(R - cur)] + ([P] if cur else [])
            assert sorted(seen) == reference_keys(attn_func, P, bc, prime)
