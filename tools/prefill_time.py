"""Time the chunked prefill (jk_prior_prefill) at several lengths of the given-token run - 384 lyric tokens of a
1b_lyrics window, and the 4096 + 384 tokens a continuation window re-primes (hop_fraction 0.5) - next to stepping the
same positions through the decode kernel, and list its kernels.  JK_WORKLOAD selects 1b_lyrics / 5b_lyrics /
small_upsampler, JK_N the number of samples."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

with contextlib.redirect_stdout(sys.stderr):
    wl = bench.SMALL if "--small" in sys.argv else bench.WORKLOADS[os.environ.get("JK_WORKLOAD", "1b_lyrics")]
    prior, _ = bench.build_prior(wl)
ca = prior.prior
n = int(os.environ.get("JK_N", "16"))
eng = ca._engine(n)
toks = torch.randint(0, ca.bins, (n, ca.input_dims), device="cuda")
yc = torch.randn(n, ca.width, device="cuda") if ca.y_cond else None
xc = torch.zeros(n, 1, ca.width, device="cuda") if ca.x_cond else None
if ca.transformer.encoder_dims:
    eng.set_encoder_kv(torch.randn(n, ca.transformer.encoder_dims, ca.width, device="cuda"))
print(f"prefill capacity {eng.prefill_capacity} positions")


def run(P):
    eng.reset(0)
    eng.prefill(n, P, tokens=toks, y_cond=yc, x_cond=xc)


lengths = [p for p in (384, 1024, 4096 + (384 if ca.input_dims > 8192 else 0)) if p <= min(eng.prefill_capacity, ca.input_dims - 1)]
for P in lengths:
    run(P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        run(P)
    e1.record()
    torch.cuda.synchronize()
    print(f"prefill of {n} x {P} positions, depth {ca.depth}: {e0.elapsed_time(e1) / 2:.2f} ms")
# the same positions stepped (what a configuration without prefill pays): 64 steps, scaled
eng.reset(0)
for _ in range(8):
    eng.step(n, tokens=toks, y_cond=yc, x_cond=xc)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(64):
    eng.step(n, tokens=toks, y_cond=yc, x_cond=xc)
e1.record()
torch.cuda.synchronize()
print(f"stepping: {e0.elapsed_time(e1) / 64:.3f} ms per position -> " +
      ", ".join(f"{P}: {e0.elapsed_time(e1) / 64 * P:.0f} ms" for P in lengths))
if lengths:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(lengths[-1])
        torch.cuda.synchronize()
    print(f"kernels of the {lengths[-1]}-position prefill:")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
