"""A few launches of the decoder-side ResConv1DBlock kernel at a level-0 shape, for `ncu -k regex:resblock`."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("JK_VARIANT"):      # A/B runs: variants/*.so
    from jukebox_b200 import _lib as _l
    _l.LIB_PATH = os.path.join(ROOT, "variants", os.environ["JK_VARIANT"] + ".so")
from jukebox_b200._lib import lib, check, ptr, stream_ptr  # noqa: E402

C = int(os.environ.get("JK_C", "64"))
T = int(os.environ.get("JK_T", "262144"))
n = int(os.environ.get("JK_N", "4"))
dil = int(os.environ.get("JK_DIL", "9"))
fn = getattr(lib(), os.environ.get("JK_FN", "jk_resblock_tc"))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, T, C, device="cuda", generator=g)
w1 = torch.randn(3, C, C, device="cuda", generator=g) / (3 * C) ** 0.5
w2 = torch.randn(C, C, device="cuda", generator=g) / C ** 0.5
b1 = torch.randn(C, device="cuda", generator=g) * 0.1
b2 = torch.randn(C, device="cuda", generator=g) * 0.1
out = torch.empty_like(x)
for _ in range(3):
    check(fn(ptr(x), ptr(out), ptr(w1), ptr(b1), ptr(w2), ptr(b2), n, T, C, dil, 1.0, stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    check(fn(ptr(x), ptr(out), ptr(w1), ptr(b1), ptr(w2), ptr(b2), n, T, C, dil, 1.0, stream_ptr()))
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1000
flops = n * T * 8 * C * C
print(f"{os.environ.get('JK_FN', 'jk_resblock_tc')} C={C} n={n} T={T} dil={dil}: {us:.1f} us, {flops / us / 1e6:.1f} TFLOP/s fp32-equivalent, "
      f"{2 * x.numel() * 4 / us / 1e3:.0f} GB/s in+out")
