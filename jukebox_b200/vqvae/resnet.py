"""Dilated residual stacks (reference: jukebox/vqvae/resnet.py:27-75), channels-last."""
import math

import torch.nn as nn

from .ops_cl import Conv1d, ReLU


class ResConv1DBlock(nn.Module):
    """x + res_scale * Conv1x1(ReLU(Conv3_dilated(ReLU(x)))) - two fused-epilogue conv launches."""

    def __init__(self, n_in, n_state, dilation=1, zero_out=False, res_scale=1.0):
        super().__init__()
        self.model = nn.Sequential(ReLU(), Conv1d(n_in, n_state, 3, 1, dilation, dilation),
                                   ReLU(), Conv1d(n_state, n_in, 1, 1, 0))
        if zero_out:
            nn.init.zeros_(self.model[-1].weight)
            nn.init.zeros_(self.model[-1].bias)
        self.res_scale = res_scale

    def forward(self, x):
        h = self.model[1](x, relu_in=True)
        return self.model[3](h, relu_in=True, res=x, scale=self.res_scale)


class Resnet1D(nn.Module):
    def __init__(self, n_in, n_depth, m_conv=1.0, dilation_growth_rate=1, dilation_cycle=None, zero_out=False,
                 res_scale=False, reverse_dilation=False, checkpoint_res=False):
        super().__init__()
        cyc = (lambda d: d) if dilation_cycle is None else (lambda d: d % dilation_cycle)
        blocks = [ResConv1DBlock(n_in, int(m_conv * n_in), dilation=dilation_growth_rate ** cyc(d),
                                 zero_out=zero_out, res_scale=1.0 if not res_scale else 1.0 / math.sqrt(n_depth))
                  for d in range(n_depth)]
        if reverse_dilation:
            blocks = blocks[::-1]
        self.checkpoint_res = checkpoint_res
        # the reference registers the same blocks under `blocks` when gradient checkpointing is
        # requested and under `model` otherwise; checkpoints depend on it (make_models.py:124-128)
        if checkpoint_res == 1:
            self.blocks = nn.ModuleList(blocks)
        else:
            self.model = nn.Sequential(*blocks)

    def forward(self, x):
        for blk in (self.blocks if self.checkpoint_res == 1 else self.model):
            x = blk(x)
        return x
