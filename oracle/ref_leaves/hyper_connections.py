"""Stand-in for hyper-connections>=0.0.10 HyperConnections; reference call
sites e2_tts.py:607, 674-677, 710-712, 818-821, 870-882, 900-939, 947. Test infrastructure only."""
from functools import partial
from random import randrange

import torch
import torch.nn.functional as F
from torch import nn
from einops import rearrange, repeat, reduce, einsum


class RMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * (self.gamma + 1)


class Residual(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, residuals):
        return residuals, (lambda out: out + residuals)


class HyperConnections(nn.Module):
    def __init__(self, num_residual_streams, *, dim, layer_index=None):
        super().__init__()
        S = num_residual_streams
        self.num_residual_streams = S
        self.norm = RMSNorm(dim)
        idx = (layer_index if layer_index is not None else randrange(S)) % S
        self.static_beta = nn.Parameter(torch.ones(S))
        a0 = torch.zeros(S, 1)
        a0[idx, 0] = 1.
        self.static_alpha = nn.Parameter(torch.cat([a0, torch.eye(S)], dim=1))
        self.dynamic_alpha_fn = nn.Parameter(torch.zeros(dim, S + 1))
        self.dynamic_alpha_scale = nn.Parameter(torch.ones(()) * 1e-2)
        self.dynamic_beta_fn = nn.Parameter(torch.zeros(dim))
        self.dynamic_beta_scale = nn.Parameter(torch.ones(()) * 1e-2)

    @classmethod
    def get_init_and_expand_reduce_stream_functions(cls, num_streams, disable=False):
        init = partial(cls if not disable else Residual, num_streams)
        if disable:
            return init, (lambda t: t), (lambda t: t)
        expand = lambda t: repeat(t, 'b ... -> (b s) ...', s=num_streams)
        red = lambda t: reduce(t, '(b s) ... -> b ...', 'sum', s=num_streams)
        return init, expand, red

    def forward(self, residuals):
        S = self.num_residual_streams
        r = rearrange(residuals, '(b s) n d -> b n s d', s=S)
        normed = self.norm(r)
        alpha = torch.tanh(normed @ self.dynamic_alpha_fn) * self.dynamic_alpha_scale + self.static_alpha
        beta = torch.tanh(normed @ self.dynamic_beta_fn) * self.dynamic_beta_scale + self.static_beta
        mix = einsum(alpha, r, 'b n s t, b n s d -> b n t d')
        branch_input, rest = mix[..., 0, :], mix[..., 1:, :]

        def add_residual(branch_out):
            out = branch_out[..., None, :] * beta[..., None] + rest
            return rearrange(out, 'b n s d -> (b s) n d')

        return branch_input, add_residual
