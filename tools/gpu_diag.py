"""GPU diagnostic (not a test): depth-1 transformers per attention pattern, comparing the engine's
fp16 intermediates (qkv, attention out, x1, gelu, h) with the oracle after every token."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import transformer_np as O          # noqa: E402
from oracle.synth import synth_state_dict         # noqa: E402
from jukebox_b200.transformer.transformer import Transformer   # noqa: E402


def run(attn_order, n_in, heads, n_ctx, blocks, bs, steps, enc_dims=0, prime_len=None, depth=1):
    tr = Transformer(n_in, n_ctx, heads, depth, mask=True, attn_order=attn_order, blocks=blocks,
                     encoder_dims=enc_dims, prime_len=prime_len)
    named = [(k, tuple(v.shape)) for k, v in tr.state_dict().items()]
    sd = synth_state_dict(named, 7)
    tr.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tr = tr.cuda().eval()
    orc = O.TransformerOracle(sd, n_in, n_ctx, heads, depth, attn_order, blocks, enc_dims, prime_len)
    rng = np.random.RandomState(0)
    x = rng.standard_normal((bs, steps, n_in)).astype(np.float32)
    enc = rng.standard_normal((bs, enc_dims, n_in)).astype(np.float32) if enc_dims else None
    worst = 0.0
    for i in range(steps):
        ref = orc.step(x[:, i], enc, True)
        with torch.no_grad():
            y = tr(torch.from_numpy(x[:, i:i + 1]).cuda(), encoder_kv=None if enc is None else torch.from_numpy(enc).cuda(),
                   sample=True, fp16=True)[:, 0].cpu().numpy()
        err = np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-9)
        worst = max(worst, err)
        if err > 1e-3 and i < 4:
            eng = tr._engine
            print(f"   step {i}: rel err {err:.3e}  |ref| {np.abs(ref).max():.3f}  |y| {np.abs(y).max():.3f}")
            for which, nm in enumerate(["h", "qkv", "a", "x1", "g"]):
                buf = eng.debug_buffer(which).float().cpu().numpy()
                print(f"      {nm}: finite={np.isfinite(buf).all()} absmax={np.abs(buf).max():.4f} first={buf[:4]}")
    print(f"attn_funcs={[l.attn_func for l in tr._attn_mods]} n_in={n_in} heads={heads} bs={bs} steps={steps}: worst rel err {worst:.3e}")
    return worst


if __name__ == "__main__":
    torch.manual_seed(0)
    print("SMs:", torch.cuda.get_device_properties(0).multi_processor_count)
    run(0, 64, 2, 48, 4, 2, 12)                       # dense
    run(0, 256, 2, 48, 4, 16, 12)                     # dense, bs 16, dh 32
    run(1, 256, 2, 48, 4, 3, 20, depth=1)             # block only
    run(2, 256, 2, 48, 4, 3, 30, depth=3)             # block, transpose, prev
    run(6, 128, 2, 48, 4, 2, 20, enc_dims=10, depth=4)
    run(12, 128, 2, 96, 8, 2, 40, prime_len=12, depth=16)
    run(0, 2048, 2, 600, 4, 16, 300, depth=1)          # 1b width, dense rows > 256 -> split-KV path
    run(2, 1024, 1, 512, 64, 16, 40, depth=3)          # upsampler-like dh 256
