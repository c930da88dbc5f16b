"""Minimal driver for ncu: build the 1b_lyrics prior (or --small), run a few decode-step launches at a
given position with teacher-forced tokens.  Not a benchmark (numbers under a profiler are never bench values)."""
import argparse
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--small", action="store_true")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--pos", type=int, default=4000)
args = ap.parse_args()
with contextlib.redirect_stdout(sys.stderr):
    prior, _ = bench.build_prior(bench.SMALL if args.small else bench.WORKLOADS['1b_lyrics'])
ca = prior.prior
n = 16
eng = ca._engine(n)
L = ca.input_dims
toks = torch.randint(0, ca.bins, (n, L), device="cuda")
lbuf = torch.empty(n, ca.bins, device="cuda")
yc = torch.randn(n, ca.width, device="cuda")
xc = torch.zeros(n, 1, ca.width, device="cuda")
lb = None       # x_cond . x_out^T: the logit bias SamplingWindow computes once per window (tensor-core logits product)
if ca.add_cond_after_transformer and eng.has_logits_gemm:
    from jukebox_b200.transformer import f32 as _f32
    lb = _f32.linear_nk(xc.reshape(n, ca.width), ca.x_out.weight).view(n, 1, ca.bins)
eng.reset(min(args.pos, L - args.steps - 1))
for _ in range(args.steps):
    eng.step(n, tokens=toks, y_cond=yc, x_cond=xc, logits=lbuf, logit_bias=lb)
torch.cuda.synchronize()
print("done", eng.position)
