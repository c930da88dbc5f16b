"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot run these shapes in seconds):

* 1b_lyrics top prior (72 layers, width 2048, n_ctx 8192 + 384 lyric tokens, 16 samples): chunked prefill ==
  stepping (the reference's check_chunks property), batch independence of a sample's logits, seeded sampling
  reproducibility through the public API;
* 3-level VQ-VAE at sample_length 1 048 576: quantise(dequantise(z)) == z bit-exact at every level, and decoding
  is local - a window of codes decodes to the same audio as the full clip, away from the window's borders.
Weights are synthetic (same generator as bench.py); nothing here reads /root/reference."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from golden_util import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prior_1b():
    import contextlib
    import bench
    with contextlib.redirect_stdout(sys.stderr):
        p, hps = bench.build_prior(bench.WORKLOADS["1b_lyrics"], seed=0)
        p._bench_hps = hps
    yield p
    del p
    torch.cuda.empty_cache()


def _logits_after(ca, n, tokens, yc, xc, P, K, use_prefill):
    ca.transformer.del_cache()
    eng = ca._engine(16)
    eng.reset(0)
    if use_prefill:
        eng.prefill(n, P, tokens=tokens, y_cond=yc, x_cond=xc)
    else:
        for _ in range(P):
            eng.step(n, tokens=tokens, y_cond=yc, x_cond=xc)
    out = torch.empty(n, K, ca.bins, device="cuda")
    lbuf = torch.empty(n, ca.bins, device="cuda")
    for k in range(K):
        eng.step(n, tokens=tokens, y_cond=yc, x_cond=xc, logits=lbuf)
        out[:, k] = lbuf
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_1b_lyrics_prefill_equals_stepping_and_batch_independence(prior_1b):
    ca = prior_1b.prior
    assert ca.transformer.n_depth == 72 and ca.width == 2048 and ca.input_dims == 8576
    n, P, K = 16, 384, 3
    g = torch.Generator().manual_seed(0)
    tokens = torch.randint(0, 79, (n, ca.input_dims), generator=g).cuda()      # lyric vocabulary for the prime
    yc = torch.randn(n, ca.width, generator=g).cuda()
    xc = torch.zeros(n, 1, ca.width, device="cuda")
    a = _logits_after(ca, n, tokens, yc, xc, P, K, use_prefill=False)
    b = _logits_after(ca, n, tokens, yc, xc, P, K, use_prefill=True)
    e = rel_err(b, a)
    print(f"1b_lyrics: prefill(384) vs 384 decode steps, logits of positions 384..386: {e:.2e}")
    assert np.isfinite(a).all() and np.isfinite(b).all() and e < 5e-3
    # a sample's logits do not depend on who else is in the batch (rows are independent end to end; only the
    # split-KV partition, hence the fp32 merge order, changes with the batch size)
    r = 5
    c = _logits_after(ca, 1, tokens[r:r + 1].contiguous(), yc[r:r + 1].contiguous(), xc[r:r + 1].contiguous(), P, K, True)
    e1 = rel_err(c[0], b[r])
    print(f"1b_lyrics: sample {r} alone vs in a batch of 16: {e1:.2e}")
    assert e1 < 2e-3


def test_1b_lyrics_seeded_sampling_is_reproducible(prior_1b):
    import bench
    prior = prior_1b
    n = 4
    y = bench.make_labels(prior, prior._bench_hps, n, seed=7).cuda()
    outs = []
    for _ in range(2):
        torch.manual_seed(123)
        z = prior.sample(n_samples=n, z=None, z_conds=None, y=y, fp16=True, temp=0.99, chunk_size=32, sample_tokens=24)
        outs.append(z.cpu())
    assert outs[0].shape == (n, 24) and torch.equal(outs[0], outs[1])
    assert int(outs[0].min()) >= 0 and int(outs[0].max()) < 2048


def test_vqvae_full_length_properties():
    import contextlib
    import bench
    from jukebox_b200.hparams import setup_hparams
    from jukebox_b200.make_models import make_vqvae
    T = 1048576
    with contextlib.redirect_stdout(sys.stderr), torch.device("cuda"):
        vq = make_vqvae(setup_hparams("vqvae", dict(sample_length=T, restore_vqvae="")), "cuda")
    bench.synth_fill(vq, 5)
    g = torch.Generator(device="cuda").manual_seed(1)
    for blk in vq.bottleneck.level_blocks:
        blk.k.normal_(generator=g)
    n = 2
    zs = [torch.randint(0, vq.l_bins, (n, T // int(h)), generator=g, device="cuda") for h in vq.hop_lengths]
    # codebook vectors are their own nearest neighbours: quantise(dequantise(z)) == z, every level, every position
    for lvl, (blk, z) in enumerate(zip(vq.bottleneck.level_blocks, zs)):
        back = blk.encode(blk.decode(z))
        assert torch.equal(back, z), f"level {lvl}: {(back != z).sum().item()} of {z.numel()} codes changed"
    # locality: decoding codes [a, b) of the top level gives the audio of the full decode away from the borders
    lvl = 2
    hop = int(vq.hop_lengths[lvl])
    full = vq.decode(zs[lvl:], start_level=lvl, bs_chunks=n)             # [n, T, 1]
    a, b = 2048, 2048 + 1024
    part = vq.decode([zs[lvl][:, a:b].contiguous()], start_level=lvl, bs_chunks=n)
    assert full.shape[1] == T and part.shape[1] == (b - a) * hop
    margin = 96 * hop                                                      # > receptive field of the level-2 decoder
    x0 = full[:, a * hop + margin:b * hop - margin]
    x1 = part[:, margin:(b - a) * hop - margin]
    d = float((x0 - x1).abs().max() / x0.abs().max())
    print(f"vqvae level {lvl}: window decode vs full decode, interior: {d:.2e}")
    assert d < 1e-5
