"""Golden vectors at BASELINE geometry, from the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_fullsize [tag ...]

The tiny fixtures of make_golden.py never reach the production code paths of the decode kernel
(head_dim 256 / 150, 47-row K/V tiles, split-KV merge, block_ctx 134, _prime_len 448, transposed layout
at p >> block_ctx, K-split GEMM groups).  These fixtures do:

  full1b_o12 : 1b_lyrics geometry (hparams.py:165-188 with n_ctx 8192 -> 8576 positions): width 2048, 2 heads
               (head_dim 256), blocks 64 (block_ctx 134), prime_len 384 (_prime_len 448), attn_order 12, depth 16
               (block / transpose / prev x5 + the prime layer at index 15)
  full1b_o9  : same geometry, attn_order 9, depth 4 (block, transpose, prev, DENSE) - the dense layer of
               1b_lyrics (index 47) at its real length
  full5b_o6  : 5b_lyrics geometry (hparams.py:127-156): width 4800, 8 heads (head_dim 150), n_ctx 8192, blocks 128
               (block_ctx 64), encoder_dims 512, fp16 Conv1D params (make_models.py:174-177); attn_order 6, depth 4
               (block, transpose, prev, ENC-DEC) instead of order 10's 19 layers to the first enc-dec layer
  fullup_o2  : released-upsampler geometry (hparams.py:68-101): width 1920, 1 head (head_dim 480), n_ctx 8192,
               blocks 128, attn_order 2, depth 3

How the reference is driven: exactly as ConditionalAutoregressive2D.primed_sample drives its transformer
(prior/autoregressive.py:300-338): `transformer(x_chunk, sample=True, fp16=...)` on chunks of given inputs
to fill the caches, and single-position calls (the sample loop body, :222-229) at the probe positions,
whose outputs are stored.  Inputs and weights are pure functions of (name, shape, seed) (oracle/synth.py),
so the fixture holds only the probe outputs (fp16 path and fp32 path).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference  # noqa: E402
from oracle.synth import synth_state_dict, synth_tensor     # noqa: E402

load_reference()
import torch as t                              # noqa: E402

FIXTURES = {
    "full1b_o12": dict(n_in=2048, n_ctx=8576, n_head=2, n_depth=16, attn_order=12, blocks=64, prime_len=384,
                       encoder_dims=0, fp16_params=False, bs=2, seed=21,
                       probes=[0, 1, 2, 133, 134, 135, 267, 268, 269, 383, 384, 385, 447, 448, 449, 4000, 8575]),
    "full1b_o9": dict(n_in=2048, n_ctx=8576, n_head=2, n_depth=4, attn_order=9, blocks=64, prime_len=None,
                      encoder_dims=0, fp16_params=False, bs=2, seed=22,
                      probes=[0, 1, 46, 47, 48, 94, 95, 134, 135, 1000, 4000, 8575]),
    "full5b_o6": dict(n_in=4800, n_ctx=8192, n_head=8, n_depth=4, attn_order=6, blocks=128, prime_len=None,
                      encoder_dims=512, fp16_params=True, bs=2, seed=23,
                      probes=[0, 1, 2, 63, 64, 65, 127, 128, 129, 4000, 8191]),
    "fullup_o2": dict(n_in=1920, n_ctx=8192, n_head=1, n_depth=3, attn_order=2, blocks=128, prime_len=None,
                      encoder_dims=0, fp16_params=False, bs=2, seed=24,
                      probes=[0, 1, 63, 64, 65, 128, 4000, 8191]),
}
CHUNK = 512


def fullsize_inputs(cfg):
    """x [bs, n_ctx, n_in] (and encoder_kv [bs, encoder_dims, n_in]); shared with tests/test_gpu_fullsize_golden.py"""
    x = synth_tensor("input.x", (cfg["bs"], cfg["n_ctx"], cfg["n_in"]), cfg["seed"])
    enc = None
    if cfg["encoder_dims"]:
        enc = synth_tensor("input.encoder_kv", (cfg["bs"], cfg["encoder_dims"], cfg["n_in"]), cfg["seed"])
    return x, enc


def run(tag):
    from jukebox.transformer.transformer import Transformer
    from jukebox.transformer.ops import _convert_conv_weights_to_fp16
    cfg = FIXTURES[tag]
    tr = Transformer(cfg["n_in"], cfg["n_ctx"], cfg["n_head"], cfg["n_depth"], mask=True,
                     attn_order=cfg["attn_order"], blocks=cfg["blocks"], encoder_dims=cfg["encoder_dims"],
                     prime_len=cfg["prime_len"])
    tr.eval()
    sd = tr.state_dict()
    named = [(k, tuple(v.shape)) for k, v in sd.items()]
    new = synth_state_dict(named, cfg["seed"])
    tr.load_state_dict({k: t.from_numpy(v) for k, v in new.items()})
    if cfg["fp16_params"]:
        tr.apply(_convert_conv_weights_to_fp16)
    x_np, enc_np = fullsize_inputs(cfg)
    x = t.from_numpy(x_np)
    enc = t.from_numpy(enc_np) if enc_np is not None else None
    outs = {}
    with t.no_grad():
        # third pass "y16_alt": the SAME reference fp16 path with another chunking (96 instead of 512 positions
        # per prefill chunk) and another BLAS thread count - every rounding point identical, only fp32
        # summation order and the chunked-vs-stepped split of the work differ.  |y16_alt - y16| is the
        # reference's own order noise; the GPU tests compare their error with it.
        for fp16, chunk, threads, key in ((True, CHUNK, 8, "y16"), (False, CHUNK, 8, "y32"), (True, 96, 3, "y16_alt")):
            t.set_num_threads(threads)
            t0 = time.time()
            tr.del_cache()
            cur, ys = 0, []
            for p in cfg["probes"]:
                while cur < p:                    # given positions, chunked (primed_sample :300-338)
                    c = min(chunk, p - cur)
                    tr(x[:, cur:cur + c].contiguous(), encoder_kv=enc, sample=True, fp16=fp16)
                    cur += c
                    tr.check_cache(cfg["bs"], cur, fp16)
                y = tr(x[:, p:p + 1].contiguous(), encoder_kv=enc, sample=True, fp16=fp16)   # the sample loop body
                cur += 1
                tr.check_cache(cfg["bs"], cur, fp16)
                ys.append(y[:, 0].float().clone())
            outs[key] = t.stack(ys, 1).numpy()          # [bs, n_probes, n_in]
            print(f"{tag}: {key} {time.time() - t0:.1f} s", flush=True)
        tr.del_cache()
    meta = dict(cfg)
    meta["attn_funcs"] = [l.attn.attn_func for l in tr._attn_mods]
    meta["chunk"] = CHUNK
    path = os.path.join(GOLDEN, tag + ".npz")
    np.savez_compressed(path, cfg=json.dumps(meta), names=json.dumps([[n, list(s), None] for n, s in named]), **outs)
    rel = float(np.abs(outs["y16"] - outs["y32"]).max() / np.abs(outs["y32"]).max())
    noise = float(np.abs(outs["y16"] - outs["y16_alt"]).max() / np.abs(outs["y16"]).max())
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB); reference fp16 vs fp32 rel {rel:.2e}, "
          f"reference fp16 order noise (chunk 512 / 8 threads vs chunk 96 / 3 threads) {noise:.2e}, "
          f"max|y| {np.abs(outs['y32']).max():.2f}", flush=True)


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    for tag in (sys.argv[1:] or list(FIXTURES)):
        run(tag)
