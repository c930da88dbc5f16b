"""Kernel breakdown of one 3-level VQ-VAE decode step (bench.py --workload vqvae_decode shapes)."""
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from jukebox_b200.hparams import setup_hparams  # noqa: E402
from jukebox_b200.make_models import make_vqvae  # noqa: E402

T = 1048576
n = 4
with contextlib.redirect_stdout(sys.stderr), torch.device("cuda"):
    vq = make_vqvae(setup_hparams("vqvae", dict(sample_length=T, restore_vqvae="")), "cuda")
bench.synth_fill(vq, 5)
zs = [torch.randint(0, vq.l_bins, (n, T // int(h)), device="cuda") for h in vq.hop_lengths]


def step():
    return [vq.decode(zs[l:], start_level=l, bs_chunks=n) for l in range(vq.levels)]


step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
