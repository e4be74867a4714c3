"""Restated x-transformers leaves (A.1-A.4). Parameter names follow upstream so that state_dict keys
match what a real install would produce. Test infrastructure only; parity for these leaves is
unpinned (no upstream install available offline)."""
from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import nn
from einops import rearrange

Intermediates = namedtuple('Intermediates', ['values'])


class RMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * self.g


class AdaptiveRMSNorm(nn.Module):
    def __init__(self, dim, dim_condition=None):
        super().__init__()
        self.scale = dim ** 0.5
        self.to_gamma = nn.Linear(dim_condition or dim, dim, bias=False)
        nn.init.zeros_(self.to_gamma.weight)

    def forward(self, x, *, condition):
        if condition.ndim == 2:
            condition = rearrange(condition, 'b d -> b 1 d')
        return F.normalize(x, dim=-1) * self.scale * (self.to_gamma(condition) + 1.)


class GLU(nn.Module):
    def __init__(self, dim_in, dim_out, activation):
        super().__init__()
        self.act = activation
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * self.act(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, glu=False, dropout=0.):
        super().__init__()
        assert glu
        inner = int(dim * mult)
        self.ff = nn.Sequential(GLU(dim, inner, nn.GELU()), nn.Dropout(dropout), nn.Linear(inner, dim))

    def forward(self, x):
        return self.ff(x)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, base=10000):
        super().__init__()
        self.register_buffer('inv_freq', 1. / (base ** (torch.arange(0, dim, 2).float() / dim)))

    def forward_from_seq_len(self, seq_len):
        t = torch.arange(seq_len, device=self.inv_freq.device)
        return self.forward(t)

    def forward(self, t):
        if t.ndim == 1:
            t = t[None]
        freqs = t.type_as(self.inv_freq)[..., None] * self.inv_freq
        freqs = torch.stack((freqs, freqs), dim=-1)
        freqs = rearrange(freqs, '... d r -> ... (d r)')
        return freqs, 1.


def rotate_half(x):
    x = rearrange(x, '... (d r) -> ... d r', r=2)
    x1, x2 = x.unbind(dim=-1)
    return rearrange(torch.stack((-x2, x1), dim=-1), '... d r -> ... (d r)')


def apply_rotary_pos_emb(t, freqs, scale=1.):
    rot_dim, seq_len, orig_dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    if t.ndim == 4 and freqs.ndim == 3:
        freqs = rearrange(freqs, 'b n d -> b 1 n d')
    t, t_un = t[..., :rot_dim], t[..., rot_dim:]
    t = (t * freqs.cos() * scale) + (rotate_half(t) * freqs.sin() * scale)
    return torch.cat((t, t_un), dim=-1).type(orig_dtype)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0., learned_value_residual_mix=False,
                 laser=False, laser_softclamp_value=15., gate_value_heads=False, softclamp_logits=False,
                 logit_softclamp_value=50.):
        super().__init__()
        assert not laser
        self.heads, self.scale = heads, dim_head ** -0.5
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_v_head_gate = None
        if gate_value_heads:
            self.to_v_head_gate = nn.Linear(dim, heads)
            nn.init.constant_(self.to_v_head_gate.weight, 0)
            nn.init.constant_(self.to_v_head_gate.bias, 10)
        self.to_value_residual_mix = None
        if learned_value_residual_mix:
            self.to_value_residual_mix = nn.Sequential(nn.Linear(dim, heads), nn.Sigmoid())
        self.softclamp_logits, self.logit_softclamp_value = softclamp_logits, logit_softclamp_value
        self.attn_dropout = nn.Dropout(dropout)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, mask=None, rotary_pos_emb=None, value_residual=None, return_intermediates=False):
        h = self.heads
        q, k, v = (rearrange(f(x), 'b n (h d) -> b h n d', h=h) for f in (self.to_q, self.to_k, self.to_v))
        orig_values = v
        if value_residual is not None:
            mix = rearrange(self.to_value_residual_mix(x), 'b n h -> b h n 1')
            v = v * mix + value_residual * (1. - mix)
        if rotary_pos_emb is not None:
            freqs, _ = rotary_pos_emb
            q, k = apply_rotary_pos_emb(q, freqs), apply_rotary_pos_emb(k, freqs)
        sim = torch.einsum('bhid,bhjd->bhij', q, k) * self.scale
        if self.softclamp_logits:
            c = self.logit_softclamp_value
            sim = (sim / c).tanh() * c
        if mask is not None:
            sim = sim.masked_fill(~mask[:, None, None, :], -torch.finfo(sim.dtype).max)
        attn = F.softmax(sim, dim=-1, dtype=torch.float32).type(sim.dtype)
        attn = self.attn_dropout(attn)
        out = torch.einsum('bhij,bhjd->bhid', attn, v)
        if self.to_v_head_gate is not None:
            gate = self.to_v_head_gate(x).sigmoid()
            out = out * rearrange(gate, 'b n h -> b h n 1')
        out = self.to_out(rearrange(out, 'b h n d -> b n (h d)'))
        if mask is not None:
            out = torch.where(mask[..., None], out, torch.zeros((), dtype=out.dtype, device=out.device))
        if not return_intermediates:
            return out
        return out, Intermediates(values=orig_values)
