"""Average duration of the decode-step kernel (no profiling stamps) at a few positions of the 1b_lyrics window."""
import contextlib
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

with contextlib.redirect_stdout(sys.stderr):
    prior = bench.build_prior("--small" in sys.argv)
ca = prior.prior
n = 16
eng = ca._engine(n)
L = ca.input_dims
toks = torch.randint(0, ca.bins, (n, L), device="cuda")
lbuf = torch.empty(n, ca.bins, device="cuda")
yc = torch.randn(n, ca.width, device="cuda")
xc = torch.zeros(n, 1, ca.width, device="cuda")
for pos in (500, 4000, 8000):
    pos = min(pos, L - 60)
    eng.reset(pos)
    for _ in range(5):
        eng.step(n, tokens=toks, y_cond=yc, x_cond=xc, logits=lbuf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.step(n, tokens=toks, y_cond=yc, x_cond=xc, logits=lbuf)
    e1.record()
    torch.cuda.synchronize()
    print(f"decode step at position {pos}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us")
