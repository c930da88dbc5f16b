"""Time the chunked prefill of the 1b_lyrics prior (16 samples x 384 lyric positions) and list its kernels."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

small = "--small" in sys.argv
with contextlib.redirect_stdout(sys.stderr):
    prior = bench.build_prior(small)
ca = prior.prior
n = 16
eng = ca._engine(n)
P = min(384, eng.prefill_capacity)
toks = torch.randint(0, ca.bins, (n, ca.input_dims), device="cuda")
yc = torch.randn(n, ca.width, device="cuda")
xc = torch.zeros(n, 1, ca.width, device="cuda")


def run():
    eng.reset(0)
    eng.prefill(n, P, tokens=toks, y_cond=yc, x_cond=xc)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
print(f"prefill of {n} x {P} positions, depth {ca.depth}: {e0.elapsed_time(e1) / 3:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
