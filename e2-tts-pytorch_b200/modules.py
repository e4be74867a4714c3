"""Host-side mirror of the reference's model classes for the hot path: same constructor / forward surface and
the same state_dict keys as /root/reference/e2_tts_pytorch/e2_tts.py, with every forward
routed through the sm_100a kernels of libb200e2tts.so (ops.py). The nn.Modules below only HOLD parameters in
the reference's layout; the arithmetic lives in the CUDA library.
"""
from __future__ import annotations

import ctypes
import math
import random as pyrandom
from collections import namedtuple
from functools import partial
from random import randrange
from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Module, ModuleList

from . import lib, ops

LossBreakdown = namedtuple('LossBreakdown', ['flow', 'velocity_consistency'])  # e2_tts.py:71
E2TTSReturn = namedtuple('E2TTS', ['loss', 'cond', 'pred_flow', 'pred_data', 'loss_breakdown'])  # e2_tts.py:73

BF16, F32 = torch.bfloat16, torch.float32
SOFTCLAMP = 50.0  # x-transformers logit_softclamp_value default (A.4)
# text sub-blocks of layer i+1 overlap the audio sub-blocks of layer i on a second CUDA stream (Transformer._run_layers);
# B200_TWO_STREAM=0 serialises them on the current stream (developer A/B switch)
import os as _os
TWO_STREAM = _os.environ.get('B200_TWO_STREAM', '1') != '0'
# depth connection of a sub-block fused into the next sub-block's width connection (ops.HcDepthWidth); B200_FUSE_HC=0: separate kernels
FUSE_HC = _os.environ.get('B200_FUSE_HC', '1') != '0'
_SIDE_STREAMS = {}


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


def _on_module_device(fn):
    """Run a public entry point with the CUDA device of the module's parameters current, so that the raw-pointer kernel launches
    (ops.py: torch.cuda.current_stream()) go to that device's stream even when the process-wide current device is another GPU.
    (Backward runs on the autograd engine's per-device threads, which set the device themselves.)"""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = next(self.parameters()).device
        if dev.type != 'cuda' or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped


def _unsupported(name, value, ref):
    raise NotImplementedError(
        f'{name}={value!r} is a non-default research switch of the reference ({ref}) for which no B200 kernel is built; '
        f'there is no fallback path (SURVEY.md §2 row 8)')


# ----------------------------------------------------------------------------------------------------------------------
# helpers kept as tiny torch ops on (B,) / (B,N) integer / bool tensors


def list_str_to_tensor(text: list[str], padding_value=-1):  # e2_tts.py:128-135 (ignores padding_value like the reference)
    rows = [torch.tensor([*bytes(t, 'UTF-8')], dtype=torch.long) for t in text]
    n = max(r.numel() for r in rows)
    return torch.stack([F.pad(r, (0, n - r.numel()), value=-1) for r in rows])


def lens_to_mask(t, length=None):  # e2_tts.py:173-182
    if not exists(length):
        length = int(t.amax())
    return torch.arange(length, device=t.device)[None, :] < t[:, None]


def mask_from_frac_lengths(seq_len, frac_lengths, max_length):  # e2_tts.py:193-210 without the .item() host sync (:189)
    lengths = (frac_lengths * seq_len).long()
    max_start = seq_len - lengths
    rand = torch.rand_like(frac_lengths)
    start = (max_start * rand).long().clamp(min=0)
    end = start + lengths
    seq = torch.arange(max_length, device=seq_len.device)
    valid = seq[None] < seq_len.max()  # positions >= max(seq_len) are padding in the reference (pad_to_length :207-208)
    return (seq[None] >= start[:, None]) & (seq[None] < end[:, None]) & valid


# ----------------------------------------------------------------------------------------------------------------------
# parameter holders (names = SURVEY Appendix B)


class RMSNorm(Module):  # A.1
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(dim))


class AdaptiveRMSNorm(Module):  # A.1
    def __init__(self, dim):
        super().__init__()
        self.to_gamma = nn.Linear(dim, dim, bias=False)
        nn.init.zeros_(self.to_gamma.weight)


class AdaLNZero(Module):  # e2_tts.py:332-351
    def __init__(self, dim, init_bias_value=-2.):
        super().__init__()
        self.to_gamma = nn.Linear(dim, dim)
        nn.init.zeros_(self.to_gamma.weight)
        nn.init.constant_(self.to_gamma.bias, init_bias_value)


class Identity(Module):
    pass


class DepthwiseConv(Module):  # e2_tts.py:295-328
    def __init__(self, dim, *, kernel_size):
        super().__init__()
        assert kernel_size % 2 == 1
        self.dw_conv1d = nn.Sequential(nn.Conv1d(dim, dim, kernel_size, groups=dim, padding=kernel_size // 2), nn.SiLU())


class Attention(Module):  # A.4
    def __init__(self, dim, heads, dim_head, learned_value_residual_mix):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_v_head_gate = nn.Linear(dim, heads)
        nn.init.constant_(self.to_v_head_gate.weight, 0)
        nn.init.constant_(self.to_v_head_gate.bias, 10)
        self.to_value_residual_mix = nn.Sequential(nn.Linear(dim, heads), nn.Sigmoid()) if learned_value_residual_mix else None
        self.to_out = nn.Linear(inner, dim, bias=False)


class _GLU(Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(Module):  # A.2
    def __init__(self, dim, mult, dropout):
        super().__init__()
        inner = int(dim * mult)
        self.ff = nn.Sequential(_GLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim))


class TextAudioCrossCondition(Module):  # e2_tts.py:486-513
    def __init__(self, dim, dim_text, cond_audio_to_text=True):
        super().__init__()
        self.text_to_audio = nn.Linear(dim_text + dim, dim, bias=False)
        nn.init.zeros_(self.text_to_audio.weight)
        self.cond_audio_to_text = cond_audio_to_text
        if cond_audio_to_text:
            self.audio_to_text = nn.Linear(dim + dim_text, dim_text, bias=False)
            nn.init.zeros_(self.audio_to_text.weight)


class _HCNorm(Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(dim))


class HyperConnections(Module):  # A.5
    def __init__(self, num_residual_streams, *, dim):
        super().__init__()
        S = num_residual_streams
        self.norm = _HCNorm(dim)
        self.static_beta = nn.Parameter(torch.ones(S))
        a0 = torch.zeros(S, 1)
        a0[randrange(S), 0] = 1.
        self.static_alpha = nn.Parameter(torch.cat([a0, torch.eye(S)], dim=1))
        self.dynamic_alpha_fn = nn.Parameter(torch.zeros(dim, S + 1))
        self.dynamic_alpha_scale = nn.Parameter(torch.ones(()) * 1e-2)
        self.dynamic_beta_fn = nn.Parameter(torch.zeros(dim))
        self.dynamic_beta_scale = nn.Parameter(torch.ones(()) * 1e-2)

    def params(self):
        return (self.norm.gamma, self.dynamic_alpha_fn, self.dynamic_alpha_scale, self.static_alpha, self.dynamic_beta_fn,
                self.dynamic_beta_scale, self.static_beta)


class RandomFourierEmbed(Module):  # e2_tts.py:355-364
    def __init__(self, dim):
        super().__init__()
        assert dim % 2 == 0
        self.register_buffer('weights', torch.randn(dim // 2))


class RotaryEmbedding(Module):  # A.3 (buffer kept for state_dict compatibility; the kernels use a cos/sin table)
    def __init__(self, dim):
        super().__init__()
        self.register_buffer('inv_freq', 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim)))


class CharacterEmbed(Module):  # e2_tts.py:390-412
    def __init__(self, dim, num_embeds=256):
        super().__init__()
        self.dim = dim
        self.embed = nn.Embedding(num_embeds + 1, dim)

    def ids(self, text, max_seq_len):
        """(b, nt) int64 with -1 padding -> (b, max_seq_len) int32 row indices (filler 0), e2_tts.py:407-410"""
        text = (text + 1)[:, :max_seq_len]
        if text.shape[1] < max_seq_len:
            text = F.pad(text, (0, max_seq_len - text.shape[1]), value=0)
        return text.to(torch.int32).contiguous()


class InterpolatedCharacterEmbed(Module):  # e2_tts.py:414-482 (E2TTS(interpolated_text=True), :1135, :1233)
    """Parameter holder with the reference's state_dict keys (`embed.weight`, `abs_pos_mlp.1.*`, `abs_pos_mlp.3.*`); the arithmetic is
    ops.InterpText. Character ids index the table directly (no +1 shift, :446) and padding (-1) is dropped per sample (:445)."""

    def __init__(self, dim, num_embeds=256):
        super().__init__()
        self.dim = dim
        self.embed = nn.Embedding(num_embeds, dim)
        self.abs_pos_mlp = nn.Sequential(nn.Identity(), nn.Linear(1, dim), nn.SiLU(), nn.Linear(dim, dim))   # index 0 = the reference's Rearrange

    @staticmethod
    def compact(text):
        """(b, nt) int64 with -1 padding anywhere -> (ids int32 (b, nt) with each row's valid characters first, in their original order;
        text_len int32 (b,)) — `one_text[one_text >= 0]` of e2_tts.py:445-446 for the whole batch."""
        valid = text >= 0
        order = torch.argsort((~valid).to(torch.int8), dim=1, stable=True)
        ids_c = torch.gather(text.clamp(min=0), 1, order).to(torch.int32).contiguous()
        return ids_c, valid.sum(dim=1).to(torch.int32)

    def embed_bf16(self, text, max_seq_len, mask=None):
        """text (b, nt) int64 with -1 padding, mask (b, n) bool | None -> bf16 [b * n, dim] (rows of masked frames are zero)."""
        B = text.shape[0]
        ids_c, text_len = self.compact(text)
        if exists(mask):
            audio_len = mask.sum(dim=1).to(torch.int32)                           # :455-457
            mask_u8 = mask.to(torch.uint8).contiguous()
        else:
            audio_len = torch.full((B,), max_seq_len, device=text.device, dtype=torch.int32)
            mask_u8 = None
        lin1, lin2 = self.abs_pos_mlp[1], self.abs_pos_mlp[3]
        return ops.InterpText.apply(ids_c, text_len, audio_len, mask_u8, self.embed.weight, lin1.weight, lin1.bias, lin2.weight, lin2.bias,
                                    B, max_seq_len)


# ----------------------------------------------------------------------------------------------------------------------
# weight packing: one kernel launch per forward refreshes every bf16 GEMM operand from the fp32 parameters


class _PackTable:
    def __init__(self):
        self.entries = []  # (param, dst, rows, cols, ld_dst, row_off, col_off, mode, out_fp32)
        self.dev_table = None
        self.ptrs = None

    def add(self, param, dst, *, ld=None, row_off=0, col_off=0, mode=0):
        p2 = param if param.dim() != 3 else param.reshape(param.shape[0], -1)
        rows, cols = (p2.shape[0], 1) if p2.dim() == 1 else p2.shape
        if ld is None:
            ld = dst.shape[-1] if dst.dim() > 1 else 1
        self.entries.append((param, dst, rows, cols, ld, row_off, col_off, mode, int(dst.dtype == F32)))

    def run(self):
        ptrs = tuple(e[0].data_ptr() for e in self.entries)
        if self.dev_table is None or ptrs != self.ptrs:
            Desc = lib.STRUCTS['b200_pack_desc']
            arr = (Desc * len(self.entries))()
            for i, (p, dst, rows, cols, ld, ro, co, mode, f32) in enumerate(self.entries):
                arr[i].src, arr[i].dst = p.data_ptr(), dst.data_ptr()
                arr[i].rows, arr[i].cols, arr[i].ld_dst, arr[i].row_off, arr[i].col_off, arr[i].mode, arr[i].out_fp32 = rows, cols, ld, ro, co, mode, f32
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self.dev_table = raw.to(self.entries[0][1].device)
            self.ptrs = ptrs
        ops.pack_weights(self.dev_table, len(self.entries))


# ----------------------------------------------------------------------------------------------------------------------


class LinearFourierEmbed(Module):
    """e2_tts.py:368-386 — parameter holder (`linear.weight`, reference state_dict key `layers.{i}.0.4.linear.weight`); the arithmetic is
    ops.FourierLinear (tcgen05 GEMM + b200_fourier_feat_*)."""

    def __init__(self, dim, p=0.5):
        super().__init__()
        assert p <= 1.
        dim_fourier = int(p * dim)
        dim_rest = dim - (dim_fourier * 2)
        self.linear = nn.Linear(dim, dim_fourier + dim_rest, bias=False)
        self.split_dims = (dim_fourier, dim_rest)


class Transformer(Module):
    """Multistream flow-matching backbone — constructor and forward signature of the reference's Transformer
    (e2_tts.py:518-952). Non-default research switches raise (no kernels, no fallback)."""

    def __init__(
        self, *, dim, dim_text=None, depth=8, heads=8, dim_head=64, ff_mult=4, text_depth=None, text_heads=None, text_dim_head=None,
        text_ff_mult=None, has_freq_axis=False, freq_heads=None, freq_dim_head=None, cond_on_time=True, abs_pos_emb=True,
        max_seq_len=8192, kernel_size=31, dropout=0.1, num_registers=32, scale_residual=False, attn_laser=False,
        attn_laser_softclamp_value=15., attn_fourier_embed_input=False, attn_fourier_embed_input_frac=0.25, num_residual_streams=4,
        attn_kwargs: dict = dict(gate_value_heads=True, softclamp_logits=True), ff_kwargs: dict = dict(),
    ):
        super().__init__()
        assert depth % 2 == 0, 'depth needs to be even'
        if has_freq_axis:
            _unsupported('has_freq_axis', has_freq_axis, 'e2_tts.py:533')
        if attn_laser:
            _unsupported('attn_laser', attn_laser, 'e2_tts.py:543')
        if dict(attn_kwargs) != dict(gate_value_heads=True, softclamp_logits=True):
            _unsupported('attn_kwargs', attn_kwargs, 'e2_tts.py:548-551')
        if dict(ff_kwargs):
            _unsupported('ff_kwargs', ff_kwargs, 'e2_tts.py:552')
        if num_residual_streams != 4:
            _unsupported('num_residual_streams', num_residual_streams, 'e2_tts.py:547')
        dim_text = default(dim_text, dim // 2)
        text_heads, text_dim_head = default(text_heads, heads), default(text_dim_head, dim_head)
        text_ff_mult, text_depth = default(text_ff_mult, ff_mult), default(text_depth, depth)
        if dim_head != 64 or text_dim_head != 64:
            _unsupported('dim_head', (dim_head, text_dim_head), 'e2_tts.py:527 (attention kernels are built for head dim 64)')
        if text_heads != heads:
            _unsupported('text_heads', text_heads, 'e2_tts.py:530')
        assert 1 <= text_depth <= depth
        assert dim % 64 == 0 and dim_text % 64 == 0 and dim <= 1024, 'kernels need dim, dim_text multiples of 64 and dim <= 1024'
        assert int(dim * ff_mult) % 64 == 0 and int(dim_text * text_ff_mult) % 64 == 0

        self.max_seq_len = max_seq_len
        self.abs_pos_emb = nn.Embedding(max_seq_len, dim) if abs_pos_emb else None
        self.dim, self.dim_text, self.depth, self.text_depth = dim, dim_text, depth, text_depth
        self.heads, self.dim_head = heads, dim_head
        self.has_freq_axis = False
        self.dropout = dropout
        self.num_streams = num_residual_streams

        self.num_registers = num_registers
        self.registers = nn.Parameter(torch.zeros(num_registers, dim))
        nn.init.normal_(self.registers, std=0.02)
        self.text_registers = nn.Parameter(torch.zeros(num_registers, dim_text))
        nn.init.normal_(self.text_registers, std=0.02)
        self.rotary_emb = RotaryEmbedding(dim_head)
        self.text_rotary_emb = RotaryEmbedding(text_dim_head)

        self.cond_on_time = cond_on_time
        norm_klass = AdaptiveRMSNorm if cond_on_time else RMSNorm
        post_klass = partial(AdaLNZero, dim) if cond_on_time else Identity
        self.time_cond_mlp = nn.Sequential(RandomFourierEmbed(dim), nn.Linear(dim + 1, dim), nn.SiLU()) if cond_on_time else Identity()

        layers, hyper_conns = [], []
        hc = partial(HyperConnections, num_residual_streams)
        for ind in range(depth):
            first, later_half, has_text = ind == 0, ind >= depth // 2, ind < text_depth
            speech = ModuleList([
                nn.Linear(dim * 2, dim, bias=False) if later_half else None,
                DepthwiseConv(dim, kernel_size=kernel_size),
                norm_klass(dim),
                Attention(dim, heads, dim_head, not first),
                LinearFourierEmbed(dim, p=attn_fourier_embed_input_frac) if attn_fourier_embed_input else nn.Identity(),   # :639
                post_klass(),
                norm_klass(dim),
                FeedForward(dim, ff_mult, dropout),
                post_klass(),
                None, None, None,
            ])
            speech_hc = ModuleList([hc(dim=dim), hc(dim=dim), hc(dim=dim), None])
            text, text_hc = None, None
            if has_text:
                text = ModuleList([
                    DepthwiseConv(dim_text, kernel_size=kernel_size),
                    RMSNorm(dim_text),
                    Attention(dim_text, text_heads, text_dim_head, not first),
                    RMSNorm(dim_text),
                    FeedForward(dim_text, text_ff_mult, dropout),
                    TextAudioCrossCondition(dim, dim_text, cond_audio_to_text=ind != text_depth - 1),
                ])
                text_hc = ModuleList([hc(dim=dim_text), hc(dim=dim_text), hc(dim=dim_text)])
            hyper_conns.append(ModuleList([speech_hc, text_hc]))
            layers.append(ModuleList([speech, text]))
        self.layers = ModuleList(layers)
        self.hyper_conns = ModuleList(hyper_conns)
        self.final_norm = RMSNorm(dim)

        self._pack = None
        self._packed = None
        self._rot = {}
        # optional int64 device word added to every dropout seed when the kernels RUN (include/b200_e2tts.h "dropout seeds"):
        # set by GraphedTrainStep, whose captured graph would otherwise replay the same dropout masks on every step
        self._seed_dev = None
        self._frozen = False   # True inside E2TTS.sample(): the bf16 operands were packed once for the whole ODE solve (weights cannot change)

    # ------------------------------------------------------------------ packed operands
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pack, self._packed, self._rot = None, None, {}
        return out

    def __deepcopy__(self, memo):  # EMA(model) deep-copies the module (trainer.py:170-174): caches are per instance
        pack, packed, rot, seed_dev = self._pack, self._packed, self._rot, self._seed_dev
        self._pack, self._packed, self._rot, self._seed_dev, self._frozen = None, None, {}, None, False
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
        finally:
            self._pack, self._packed, self._rot, self._seed_dev = pack, packed, rot, seed_dev
        return new

    def _build_pack(self):
        dev = self.registers.device
        d, dt, H = self.dim, self.dim_text, self.heads
        I = H * 64
        tab, cond_tab, packed = _PackTable(), _PackTable(), []
        e = lambda *s: torch.zeros(s, device=dev, dtype=BF16)
        cond_rows = []
        for i, (speech, text) in enumerate(self.layers):
            L = {}
            for pre, mods, din in (('a', speech, d), ('t', text, dt)):
                if mods is None:
                    continue
                attn = mods[3] if pre == 'a' else mods[2]
                ff = mods[7] if pre == 'a' else mods[4]
                has_mix = attn.to_value_residual_mix is not None
                qkv = e(3 * I + (2 if has_mix else 1) * H, din)
                for j, lin in enumerate((attn.to_q, attn.to_k, attn.to_v)):
                    tab.add(lin.weight, qkv, row_off=j * I)
                tab.add(attn.to_v_head_gate.weight, qkv, row_off=3 * I)
                if has_mix:
                    tab.add(attn.to_value_residual_mix[0].weight, qkv, row_off=3 * I + H)
                out_w = e(din, I)
                tab.add(attn.to_out.weight, out_w)
                inner = ff.ff[2].weight.shape[1]
                w1, b1 = e(2 * inner, din), torch.zeros(2 * inner, device=dev, dtype=F32)
                tab.add(ff.ff[0].proj.weight, w1, mode=1)
                tab.add(ff.ff[0].proj.bias, b1, mode=1)
                w2 = e(din, inner)
                tab.add(ff.ff[2].weight, w2)
                L[pre] = dict(qkv=qkv, out=out_w, w1=w1, b1=b1, w2=w2)
            if isinstance(speech[4], LinearFourierEmbed):
                lfe = e(*speech[4].linear.weight.shape)
                tab.add(speech[4].linear.weight, lfe)
                L['a']['lfe'] = lfe
            if speech[0] is not None:
                L['skip'] = e(d, 2 * d)
                tab.add(speech[0].weight, L['skip'])
            if text is not None:
                cc = text[5]
                stack = e(d + (dt if cc.cond_audio_to_text else 0), d + dt)
                tab.add(cc.text_to_audio.weight, stack)
                if cc.cond_audio_to_text:
                    tab.add(cc.audio_to_text.weight, stack, row_off=d)
                L['cross'] = stack
            if self.cond_on_time:
                cond_rows += [speech[2].to_gamma, speech[5].to_gamma, speech[6].to_gamma, speech[8].to_gamma]
            packed.append(L)
        cond = None
        if self.cond_on_time:
            n = len(cond_rows) * d
            W_all = torch.zeros((n, d), device=dev, dtype=F32)
            b_all = torch.zeros(n, device=dev, dtype=F32)
            for j, lin in enumerate(cond_rows):
                cond_tab.add(lin.weight, W_all, row_off=j * d)
                if lin.bias is not None:
                    cond_tab.add(lin.bias, b_all, row_off=j * d)
            cond = dict(W=W_all, b=b_all, lins=cond_rows)
        self._pack, self._packed = (tab, cond_tab), dict(layers=packed, cond=cond)

    def refresh_packed(self):
        """Re-pack the bf16 GEMM operands from the current fp32 parameters (one launch)."""
        if self._frozen and self._pack is not None:
            return
        if self._pack is None:
            self._build_pack()
        self._pack[0].run()

    def freeze_packed(self, on):
        """sample() runs 124 forwards over frozen weights (e2_tts.py:1332 @torch.no_grad, :1351 eval): pack the tensor-core operands and the
        batched to_gamma matrix ONCE instead of once per forward."""
        self._frozen = False
        if on:
            self.refresh_packed()
            if self.cond_on_time:
                self._pack[1].run()
            self._frozen = True

    def _rotary(self, Np, dev):
        if Np not in self._rot:
            self._rot[Np] = ops.rotary_table(Np, dev)
        return self._rot[Np]

    # ------------------------------------------------------------------ conditioning vectors
    def _cond_gains(self, times, batch):
        """time_cond_mlp (:621-625, 778-789) and every per-layer to_gamma projection in ONE batched launch.
        Returns a list [4 * depth] of contiguous fp32 [B, d]: (1 + gamma) for the adaptive norms, sigmoid gates for AdaLNZero."""
        if times.ndim == 0:
            times = times.expand(batch)
        times = times.to(F32).contiguous()
        four = ops.fourier_embed(times, self.time_cond_mlp[0].weights)
        lin = self.time_cond_mlp[1]
        cond = ops.SmallLinear.apply(four, lin.weight, lin.bias, 1, self.dim, False)
        c = self._packed['cond']
        tab = self._pack[1]
        weights = [l.weight for l in c['lins']]
        biases = [l.bias for l in c['lins'] if l.bias is not None]  # AdaLNZero gates sit on the odd d-wide segments
        if self._frozen:
            W, b_full = c['W'], c['b']
        else:
            W, b_full = ops.CondPack.apply(tab.run, c['W'], c['b'], self.dim, len(weights), *weights, *biases)
        gains = ops.SmallLinear.apply(cond, W, b_full, 5, self.dim, True)
        return list(gains.unbind(0))

    # ------------------------------------------------------------------ the block stack
    def _run_layers(self, xs, ts, gains, mask_u8, B, Np, seed):
        """xs bf16 [T,S,d], ts bf16 [T,S,dt] | None -> final residual streams. Layer loop of e2_tts.py:825-939."""
        P = self._packed['layers']
        H = self.heads
        cs, sn = self._rotary(Np, xs.device)
        p_drop = self.dropout if self.training else 0.0
        mbits = ops.attn_maskbits(mask_u8, B, Np, xs.device) if xs.is_cuda else None   # one key-mask bitmask for all 2 * depth attention calls
        skips = []
        v_first, tv_first = None, None
        nseed = [seed]

        def next_seed():
            nseed[0] = (nseed[0] * 6364136223846793005 + 1442695040888963407) & 0x7FFFFFFFFFFFFFFF
            return nseed[0]

        # Hyper-connection plumbing of one stream: a sub-block returns its depth connection PENDING — (residual', branch_out, beta) — and
        # the next sub-block's width connection consumes it in one fused kernel (ops.HcDepthWidth); `close` materialises the streams
        # where something else reads them (cross-conditioning, skip path, final norm).
        fuse = FUSE_HC and ops.hc_can_fuse(xs.shape[0], xs.shape[1])

        def width(res, hcm, gain, mode):
            if isinstance(res, tuple):
                return ops.HcDepthWidth.apply(*res, *hcm.params(), gain, mode, Np)
            return ops.HcWidth.apply(res, *hcm.params(), gain, mode, Np)

        def depth(rest, y, beta):
            return (rest, y, beta) if fuse else ops.HcDepth.apply(rest, y, beta)

        def close(res):
            return ops.HcDepth.apply(*res) if isinstance(res, tuple) else res

        def sub_conv(res, hcm, conv):
            br, rest, beta = width(res, hcm, None, 0)
            y = ops.DwConv.apply(br, conv.dw_conv1d[0].weight, conv.dw_conv1d[0].bias, mask_u8, B, Np)
            return depth(rest, y, beta)

        def sub_attn(res, hcm, gain, mode, attn, pk, vf, colscale, lfe=None):
            br, rest, beta = width(res, hcm, gain, mode)
            if lfe is not None:   # attn_input_fourier_embed (:909): between the attention norm (fused into the width kernel) and the attention
                br = ops.FourierLinear.apply(br, lfe.linear.weight, pk['lfe'], *lfe.split_dims)
            mix = attn.to_value_residual_mix
            og, v = ops.Attention.apply(br, attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_v_head_gate.weight,
                                        attn.to_v_head_gate.bias, mix[0].weight if mix is not None else None,
                                        mix[0].bias if mix is not None else None, vf if mix is not None else None,
                                        pk['qkv'], cs, sn, mask_u8, B, Np, H, p_drop, next_seed(), SOFTCLAMP, self._seed_dev, mbits)
            y = ops.OutProj.apply(og, attn.to_out.weight, pk['out'], colscale, mask_u8, B, Np)
            return depth(rest, y, beta), (v if vf is None else vf)

        def sub_ff(res, hcm, gain, mode, ff, pk, colscale):
            br, rest, beta = width(res, hcm, gain, mode)
            y = ops.FeedForward.apply(br, ff.ff[0].proj.weight, ff.ff[0].proj.bias, ff.ff[2].weight, ff.ff[2].bias,
                                      pk['w1'], pk['b1'], pk['w2'], colscale, B, Np, p_drop, next_seed(), self._seed_dev)
            return close(depth(rest, y, beta))

        def text_block(i, ts, tvf):  # the three text sub-blocks of layer i (:853-882)
            text, thc, pk = self.layers[i][1], self.hyper_conns[i][1], P[i]['t']
            ts = sub_conv(ts, thc[0], text[0])
            ts, tvf = sub_attn(ts, thc[1], text[1].g, 1, text[2], pk, tvf, None)
            return sub_ff(ts, thc[2], text[3].g, 1, text[4], pk, None), tvf

        # The text sub-blocks of layer i+1 depend on the cross-conditioning of layer i only, not on layer i's audio sub-blocks
        # (:853-939): enqueue them on a second stream so that the half-width text kernels (K = dt GEMMs, D = dt token kernels)
        # fill the launch gaps and tile-quantisation tails of the audio kernels instead of serialising with them. The autograd
        # engine replays the same fork/join in backward (each node runs on its forward stream). Tensors that cross streams are
        # registered with the caching allocator (record_stream).
        has_text = lambda i: ts is not None and i < len(self.layers) and self.layers[i][1] is not None
        two = TWO_STREAM and has_text(0) and xs.is_cuda
        if two:
            main, side = torch.cuda.current_stream(xs.device), _side_stream(xs.device)
            for t in (mask_u8, cs, sn, mbits):
                if t is not None:
                    t.record_stream(side)

            def fork_text(i, ts_in, tvf):
                side.wait_stream(main)
                ts_in.record_stream(side)
                with torch.cuda.stream(side):
                    return text_block(i, ts_in, tvf)
            ts, tv_first = fork_text(0, ts, tv_first)

        for i, ((speech, text), (shc, thc)) in enumerate(zip(self.layers, self.hyper_conns)):
            pk = P[i]
            if ts is not None and text is not None:  # :853-883
                if two:
                    main.wait_stream(side)      # join: text sub-blocks of this layer (enqueued one layer ago)
                    ts.record_stream(main)
                else:
                    ts, tv_first = text_block(i, ts, tv_first)
                cc = text[5]
                xs, ts = ops.CrossCondition.apply(xs, ts, cc.text_to_audio.weight,
                                                  cc.audio_to_text.weight if cc.cond_audio_to_text else None, pk['cross'])
                if two and has_text(i + 1):
                    ts, tv_first = fork_text(i + 1, ts, tv_first)
            if (i + 1) <= self.depth // 2:  # :887-896
                skips.append(xs)
            else:
                xs = ops.SkipProj.apply(xs, skips.pop(), speech[0].weight, pk['skip'])
            if self.cond_on_time:
                g_an, g_az, g_fn, g_fz = gains[4 * i:4 * i + 4]
                mode = 2
            else:
                g_an, g_az, g_fn, g_fz = speech[2].g, None, speech[6].g, None
                mode = 1
            xs = sub_conv(xs, shc[0], speech[1])  # :900-902
            xs, v_first = sub_attn(xs, shc[1], g_an, mode, speech[3], pk['a'], v_first, g_az,
                                   speech[4] if isinstance(speech[4], LinearFourierEmbed) else None)  # :906-916
            xs = sub_ff(xs, shc[2], g_fn, mode, speech[7], pk['a'], g_fz)  # :936-939
        assert not skips
        return xs

    def _prepare(self, B, N, mask):
        Np = N + self.num_registers
        assert N <= self.max_seq_len, f'{N} exceeds the set `max_seq_len` ({self.max_seq_len}) on Transformer'
        mask_u8 = None
        if exists(mask):
            mask_u8 = F.pad(mask, (self.num_registers, 0), value=True).to(torch.uint8).contiguous()
        return Np, mask_u8

    @_on_module_device
    def forward(self, x, times=None, mask=None, text_embed=None):
        """Reference signature (e2_tts.py:731-737): x (b, n, d) -> (b, n, d)."""
        assert x.ndim == 3, '`has_freq_axis` is not supported by the B200 build'
        assert not (exists(times) ^ self.cond_on_time), '`times` must be passed in if `cond_on_time` is set to `True` and vice versa'
        B, N, d = x.shape
        if torch.is_grad_enabled():
            ops.zero_pool.begin(x.device)
        h = ops.CastRows.apply(x.reshape(B * N, d))
        te = ops.CastRows.apply(text_embed.reshape(B * N, -1)) if exists(text_embed) else None
        y = self._forward_from_h(h, B, N, times, mask, te_bf16=te)
        return y.view(B, N, d).to(x.dtype)

    def _forward_from_h(self, h, B, N, times, mask, text_ids=None, text_embed_module=None, te_bf16=None):
        """h bf16 [B*N, d] (already projected) -> final-normed bf16 [B*N, d]."""
        self.refresh_packed()
        S, R = self.num_streams, self.num_registers
        Np, mask_u8 = self._prepare(B, N, mask)
        abs_w = self.abs_pos_emb.weight if exists(self.abs_pos_emb) else None
        xs = ops.Assemble.apply(h, abs_w, self.registers, B, N, S)
        ts = None
        if exists(text_ids):
            ts = ops.TextStem.apply(text_ids, text_embed_module.embed.weight, self.text_registers, B, N, S)
        elif exists(te_bf16):
            ts = ops.Assemble.apply(te_bf16, None, self.text_registers, B, N, S)
        gains = self._cond_gains(times, B) if self.cond_on_time else None
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (self.training and self.dropout > 0) else 0
        xs = self._run_layers(xs, ts, gains, mask_u8, B, Np, seed)
        return ops.FinalNorm.apply(xs, self.final_norm.g, B, N, R)


# ----------------------------------------------------------------------------------------------------------------------
# MelSpec (e2_tts.py:248-290): parameter/buffer holder with torchaudio-compatible buffer names (Appendix B)


class _Spectrogram(Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer('window', torch.hann_window(n_fft, periodic=True))


class _MelScale(Module):
    def __init__(self, n_freqs, n_mels, sample_rate):
        super().__init__()
        self.register_buffer('fb', mel_filterbank(n_freqs, n_mels, sample_rate))


class _MelSTFT(Module):
    def __init__(self, n_fft, n_mels, sample_rate):
        super().__init__()
        self.spectrogram = _Spectrogram(n_fft)
        self.mel_scale = _MelScale(n_fft // 2 + 1, n_mels, sample_rate)


def mel_filterbank(n_freqs, n_mels, sample_rate, f_min=0.0, f_max=None):
    """HTK mel filterbank, norm=None — what torchaudio.transforms.MelSpectrogram builds for the reference (:265-275)."""
    f_max = f_max if f_max is not None else sample_rate / 2
    hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz2mel(f_min), hz2mel(f_max), n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    return torch.clamp(torch.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)


class MelSpec(Module):
    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=100, sampling_rate=24_000,
                 normalize=False, power=1, norm=None, center=True):
        super().__init__()
        if (win_length, normalize, power, norm, center) != (filter_length, False, 1, None, True):
            _unsupported('mel_spec_kwargs', dict(win_length=win_length, normalize=normalize, power=power, norm=norm, center=center),
                         'e2_tts.py:249-260')
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.n_fft, self.hop = filter_length, hop_length
        self.mel_stft = _MelSTFT(filter_length, n_mel_channels, sampling_rate)
        self.register_buffer('dummy', torch.tensor(0), persistent=False)

    def forward(self, inp):
        if inp.ndim == 3:
            inp = inp[:, 0]
        assert inp.ndim == 2
        if self.dummy.device != inp.device:
            self.to(inp.device)
        return ops.melspec(inp.to(F32).contiguous(), self.mel_stft.spectrogram.window, self.mel_stft.mel_scale.fb, self.n_fft, self.hop)

    def collate(self, waves, lens=None):
        """On-device data path: what the reference does per item on CPU workers — `MelSpec` in HFDataset.__getitem__
        (trainer.py:101-131) — and per batch in `collate_fn` (:61-82: zero-pad the mels to the longest, lengths) plus the trainer's
        `rearrange(batch['mel'], 'b d n -> b n d')` (:253), as ONE kernel launch over the ragged batch.
        waves: list of 1-D fp32 tensors at `sampling_rate` (resampling is the dataset's job) or a zero-padded [B, nw_max] tensor with
        `lens` (samples per sequence). Returns dict(mel [B, n_frames_max, n_mels] fp32, mel_lengths [B] int64) — pass as
        `model(batch['mel'], text=..., lens=batch['mel_lengths'])`."""
        if isinstance(waves, (list, tuple)):
            dev = self.dummy.device if self.dummy.device.type == 'cuda' else waves[0].device
            lens = torch.tensor([w.shape[-1] for w in waves], dtype=torch.int32)
            nmax = int(lens.max())
            padded = torch.zeros((len(waves), nmax), dtype=F32).pin_memory() if dev.type == 'cuda' and not waves[0].is_cuda else torch.zeros((len(waves), nmax), dtype=F32, device=waves[0].device)
            for i, w in enumerate(waves):
                padded[i, :w.shape[-1]] = w.reshape(-1)
            waves = padded.to(dev, non_blocking=True)
            lens = lens.to(dev, non_blocking=True)
        else:
            assert lens is not None and waves.ndim == 2
            lens = lens.to(device=waves.device, dtype=torch.int32)
        if self.dummy.device != waves.device:
            self.to(waves.device)
        mel = ops.melspec(waves.to(F32).contiguous(), self.mel_stft.spectrogram.window, self.mel_stft.mel_scale.fb, self.n_fft, self.hop,
                          wave_lens=lens.contiguous(), out_bnd=True)
        mel_lengths = 1 + lens.long() // self.hop
        n_max = int(1 + waves.shape[1] // self.hop)
        return dict(mel=mel[:, :n_max], mel_lengths=mel_lengths)


# ----------------------------------------------------------------------------------------------------------------------


class _HLGaussRegression(Module):  # A.6, regression mode
    def __init__(self, dim):
        super().__init__()
        self.to_pred = nn.Sequential(nn.Linear(dim, 1), nn.Softplus())


def _resolve_tokenizer(tokenizer, text_num_embeds):
    if callable(tokenizer):
        assert exists(text_num_embeds), '`text_num_embeds` must be given if supplying your own tokenizer encode function'
        return tokenizer, text_num_embeds
    if tokenizer == 'char_utf8':
        return list_str_to_tensor, 256
    if tokenizer == 'phoneme_en':
        _unsupported('tokenizer', tokenizer, 'e2_tts.py:141-166 (g2p_en is host-side preprocessing; pass a callable tokenizer instead)')
    raise ValueError(f'unknown tokenizer string {tokenizer}')


class DurationPredictor(Module):
    """Reference surface: e2_tts.py:956-1113."""

    def __init__(self, transformer: dict | Transformer, num_channels=None, mel_spec_kwargs: dict = dict(), char_embed_kwargs: dict = dict(),
                 text_num_embeds=None, num_freq_tokens=1, hl_gauss_loss: dict | None = None, use_regression=True,
                 tokenizer: str | Callable = 'char_utf8'):
        super().__init__()
        if num_freq_tokens != 1:
            _unsupported('num_freq_tokens', num_freq_tokens, 'e2_tts.py:965')
        if hl_gauss_loss is not None or not use_regression:
            _unsupported('hl_gauss_loss', hl_gauss_loss, 'e2_tts.py:966-967 (only the regression head is used by the reference defaults)')
        self.num_freq_tokens, self.has_freq_axis = 1, False
        if isinstance(transformer, dict):
            transformer = dict(transformer)
            transformer.setdefault('has_freq_axis', False)
            transformer = Transformer(**transformer, cond_on_time=False)
        assert not transformer.has_freq_axis
        self.mel_spec = MelSpec(**mel_spec_kwargs)
        self.num_channels = default(num_channels, self.mel_spec.n_mel_channels)
        self.transformer = transformer
        self.dim = transformer.dim
        self.proj_in = nn.Linear(self.num_channels, self.dim)
        self.tokenizer, text_num_embeds = _resolve_tokenizer(tokenizer, text_num_embeds)
        self.embed_text = CharacterEmbed(transformer.dim_text, num_embeds=text_num_embeds, **char_embed_kwargs)
        self.hl_gauss_layer = _HLGaussRegression(self.dim)
        self._wpack = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._wpack = None
        return out

    @_on_module_device
    def forward(self, x, *, text=None, lens=None, return_loss=True):
        if x.ndim == 2:  # raw wave (:1052-1055; the reference's `== self.dim` assert is a known bug, Appendix C)
            x = self.mel_spec(x).transpose(1, 2)
        x = x.to(F32).contiguous()
        B, N, C = x.shape
        dev = x.device
        if torch.is_grad_enabled() and return_loss:
            ops.zero_pool.begin(dev)
        Cp = (C + 7) // 8 * 8
        if self._wpack is None or self._wpack[0].device != dev:
            w = torch.zeros((self.dim, Cp), device=dev, dtype=BF16)
            tab = _PackTable()
            tab.add(self.proj_in.weight, w)
            self._wpack = (w, tab)
        self._wpack[1].run()
        A = ops.cast_rows(x.reshape(B * N, C), B * N, C, Cp)
        h = ops.StemLinear.apply(A, self.proj_in.weight, self.proj_in.bias, None, None, self._wpack[0])
        ids = None
        if exists(text):
            if isinstance(text, list):
                text = list_str_to_tensor(text).to(dev)  # :1067 (always the byte tokenizer, Appendix C)
                assert text.shape[0] == B
            ids = self.embed_text.ids(text, N)
        if not exists(lens):
            lens = torch.full((B,), N, device=dev)
        mask = lens_to_mask(lens, length=N)
        if return_loss:  # :1081-1086
            rand_frac_index = _rng.draw('duration_rand_frac', lambda: x.new_zeros(B).uniform_(0, 1))
            rand_index = (rand_frac_index * lens).long()
            mask = mask & (torch.arange(N, device=dev)[None] < rand_index[:, None])
        tr = self.transformer
        y = tr._forward_from_h(h, B, N, None, mask, text_ids=ids, text_embed_module=self.embed_text)
        pooled = ops.MaskedMean.apply(y, mask.to(torch.uint8).contiguous(), B, N)
        lin = self.hl_gauss_layer.to_pred[0]
        pred = ops.SmallLinear.apply(pooled, lin.weight, lin.bias, 4, 1, False).squeeze(-1)
        if not return_loss:
            return pred
        return F.mse_loss(pred, lens.float())  # (B,) scalar glue, :1111


class _RngOverride:
    """Test hook: replay the reference's random draws (x0, times, span mask, drop_text_cond, duration prefix) so that
    parity tests run all three implementations on identical (mel, text, t, noise) inputs."""

    def __init__(self):
        self.values = None

    def draw(self, name, fn):
        if self.values is not None and name in self.values:
            return self.values[name]
        return fn()


_rng = _RngOverride()


class inject_randomness:
    def __init__(self, **values):
        self.values = values

    def __enter__(self):
        _rng.values = self.values

    def __exit__(self, *a):
        _rng.values = None


class E2TTS(Module):
    """Reference surface: e2_tts.py:1115-1595 (constructor, forward, sample, transformer_with_pred_head,
    cfg_transformer_with_pred_head, device)."""

    def __init__(self, transformer: dict | Transformer = None, duration_predictor: dict | DurationPredictor | None = None,
                 odeint_kwargs: dict = dict(atol=1e-5, rtol=1e-5, method='midpoint'), cond_drop_prob=0.25, num_channels=None,
                 mel_spec_module: Module | None = None, num_freq_tokens=1, char_embed_kwargs: dict = dict(), mel_spec_kwargs: dict = dict(),
                 frac_lengths_mask: tuple[float, float] = (0.7, 1.), concat_cond=False, interpolated_text=False,
                 text_num_embeds: int | None = None, tokenizer: str | Callable = 'char_utf8', use_vocos=True,
                 pretrained_vocos_path='charactr/vocos-mel-24khz', sampling_rate: int | None = None, velocity_consistency_weight=0.):
        super().__init__()
        if num_freq_tokens != 1:
            _unsupported('num_freq_tokens', num_freq_tokens, 'e2_tts.py:1130')
        if odeint_kwargs.get('method', 'midpoint') not in ('midpoint', 'euler'):
            _unsupported('odeint_kwargs', odeint_kwargs, 'e2_tts.py:1122-1126 (fixed-grid midpoint/euler only)')
        self.num_freq_tokens, self.has_freq_axis = 1, False
        if isinstance(transformer, dict):
            transformer = dict(transformer)
            transformer.setdefault('has_freq_axis', False)
            transformer = Transformer(**transformer, cond_on_time=True)
        assert not transformer.has_freq_axis
        self.transformer = transformer
        if isinstance(duration_predictor, dict):
            duration_predictor = DurationPredictor(**duration_predictor)
        dim, dim_text = transformer.dim, transformer.dim_text
        self.dim, self.dim_text = dim, dim_text
        self.frac_lengths_mask = frac_lengths_mask
        self.duration_predictor = duration_predictor
        self.odeint_kwargs = odeint_kwargs
        self.mel_spec = default(mel_spec_module, MelSpec(**mel_spec_kwargs))
        num_channels = default(num_channels, self.mel_spec.n_mel_channels)
        self.num_channels = num_channels
        self.sampling_rate = default(sampling_rate, getattr(self.mel_spec, 'sampling_rate', None))
        self.concat_cond = concat_cond
        if concat_cond:      # :1200-1204: one Linear on cat(cond, x) instead of two summed projections
            self.proj_in = nn.Linear(num_channels * 2, dim)
        else:
            self.proj_in = nn.Linear(num_channels, dim)
            self.cond_proj_in = nn.Linear(num_channels, dim)
        self.to_pred = nn.Linear(dim, num_channels)
        self.tokenizer, text_num_embeds = _resolve_tokenizer(tokenizer, text_num_embeds)
        self.cond_drop_prob = cond_drop_prob
        text_embed_klass = InterpolatedCharacterEmbed if interpolated_text else CharacterEmbed          # :1233
        self.embed_text = text_embed_klass(dim_text, num_embeds=text_num_embeds, **char_embed_kwargs)
        self.register_buffer('zero', torch.tensor(0.), persistent=False)
        self.velocity_consistency_weight = velocity_consistency_weight
        # Vocos is a separate pretrained network fetched from the HF hub (e2_tts.py:1244): out of scope.
        self.vocos = None
        self._use_vocos_requested = use_vocos
        self._wpack = None

    @property
    def device(self):
        return next(self.parameters()).device

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._wpack = None
        return out

    def _packed(self):
        dev = self.device
        if self.transformer._frozen and self._wpack is not None and self._wpack['stem'].device == dev:
            return self._wpack
        if self._wpack is None or self._wpack['stem'].device != dev:
            C, d = self.num_channels, self.dim
            Cp = (C + 63) // 64 * 64
            stem = torch.zeros((d, 2 * Cp), device=dev, dtype=BF16)
            pred = torch.zeros((C, d), device=dev, dtype=BF16)
            tab = _PackTable()
            tab.add(self.proj_in.weight, stem)          # concat_cond: all 2C columns, matching stem_prepare's cat(cond, x) layout
            if not self.concat_cond:
                tab.add(self.cond_proj_in.weight, stem, col_off=Cp)
            tab.add(self.to_pred.weight, pred)
            self._wpack = dict(stem=stem, pred=pred, tab=tab, Cp=Cp)
        self._wpack['tab'].run()
        return self._wpack

    def _embed(self, A, B, N, times, mask, text, drop_text_cond, pk):
        if self.concat_cond:
            h = ops.StemLinear.apply(A, self.proj_in.weight, self.proj_in.bias, None, None, pk['stem'])
        else:
            h = ops.StemLinear.apply(A, self.proj_in.weight, self.proj_in.bias, self.cond_proj_in.weight, self.cond_proj_in.bias, pk['stem'])
        ids, te = None, None
        if exists(text) and not drop_text_cond:
            if isinstance(self.embed_text, InterpolatedCharacterEmbed):
                te = self.embed_text.embed_bf16(text, N, mask)      # :1283 embed_text(text, seq_len, mask = mask)
            else:
                ids = self.embed_text.ids(text, N)
        y = self.transformer._forward_from_h(h, B, N, times, mask, text_ids=ids, text_embed_module=self.embed_text, te_bf16=te)
        return y, pk

    @_on_module_device
    def transformer_with_pred_head(self, x, cond, times, mask=None, text=None, drop_text_cond=None, return_drop_text_cond=False):
        """e2_tts.py:1250-1301."""
        B, N, C = x.shape
        drop_text_cond = default(drop_text_cond, self.training and pyrandom.random() < self.cond_drop_prob)
        pk = self._packed()
        A, _ = ops.stem_prepare(B, N, C, pk['Cp'], x_in=x.to(F32).contiguous(), cond_in=cond.to(F32).contiguous(), concat=self.concat_cond)
        if not torch.is_tensor(times):
            times = torch.tensor(times, device=x.device)
        y, pk = self._embed(A, B, N, times.to(x.device), mask, text, drop_text_cond, pk)
        pred = ops.PredHead.apply(y, self.to_pred.weight, self.to_pred.bias, pk['pred']).view(B, N, C).to(x.dtype)
        if not return_drop_text_cond:
            return pred
        return pred, drop_text_cond

    def cfg_transformer_with_pred_head(self, *args, cfg_strength: float = 1., cfg_null_model=None, remove_parallel_component: bool = True,
                                       keep_parallel_frac: float = 0., **kwargs):
        """e2_tts.py:1303-1330 (CFG + APG projection in fp64, :113-124)."""
        pred = self.transformer_with_pred_head(*args, drop_text_cond=False, **kwargs)
        if cfg_strength < 1e-5:
            return pred
        null_drop = not exists(cfg_null_model)
        cfg_null_model = default(cfg_null_model, self)
        null_pred = cfg_null_model.transformer_with_pred_head(*args, drop_text_cond=null_drop, **kwargs)
        return ops.cfg_combine(pred, null_pred, float(cfg_strength), bool(remove_parallel_component), float(keep_parallel_frac))

    @torch.no_grad()
    @_on_module_device
    def sample(self, cond, *, text=None, lens=None, duration=None, steps=32, cfg_strength=1., cfg_null_model=None, max_duration=4096,
               vocoder=None, return_raw_output=None, save_to_filename=None):
        """e2_tts.py:1332-1466. Fixed-grid ODE on t = linspace(0, 1, steps) (torchdiffeq semantics)."""
        self.eval()
        if not (exists(return_raw_output) and return_raw_output) and not exists(vocoder) and (self._use_vocos_requested or exists(save_to_filename)):
            # fail BEFORE the ODE loop (124 transformer passes at the defaults), not after it
            raise NotImplementedError('Vocos decoding / audio saving needs the pretrained vocoder from the HF hub and is out of scope '
                                      '(SURVEY.md §2 row 10): call sample(..., return_raw_output=True), pass `vocoder=`, or build E2TTS(use_vocos=False)')
        if cond.ndim == 2:
            cond = self.mel_spec(cond).transpose(1, 2)
            assert cond.shape[-1] == self.num_channels
        cond = cond.to(F32)
        B, cond_seq_len, dev = *cond.shape[:2], cond.device
        if not exists(lens):
            lens = torch.full((B,), cond_seq_len, device=dev, dtype=torch.long)
        if isinstance(text, list):
            text = self.tokenizer(text).to(dev)
            assert text.shape[0] == B
        if exists(text):
            lens = torch.maximum((text != -1).sum(dim=-1), lens)
        cond_mask = lens_to_mask(lens)
        if exists(duration):
            if isinstance(duration, int):
                duration = torch.full((B,), duration, device=dev, dtype=torch.long)
        elif exists(self.duration_predictor):
            duration = self.duration_predictor(cond, text=text, lens=lens, return_loss=False).long()
        duration = torch.maximum(lens + 1, duration).clamp(max=max_duration)
        assert duration.shape[0] == B
        md = int(duration.amax())
        cond = F.pad(cond, (0, 0, 0, md - cond_seq_len), value=0.)
        cond_mask = F.pad(cond_mask, (0, md - cond_mask.shape[-1]), value=False)[..., None]
        mask = lens_to_mask(duration)
        step_cond = torch.where(cond_mask, cond, torch.zeros_like(cond))  # step-invariant: hoisted out of fn (:1404)

        def fn(t, x):
            return self.cfg_transformer_with_pred_head(x, step_cond, times=t, text=text, mask=mask, cfg_strength=cfg_strength,
                                                       cfg_null_model=cfg_null_model)

        y = _rng.draw('y0', lambda: torch.randn_like(cond))
        ts_host = torch.linspace(0, 1, steps)          # fixed grid (torchdiffeq semantics); step sizes stay host floats: no device sync per step
        ts = ts_host.to(dev)
        method = self.odeint_kwargs.get('method', 'midpoint')
        frozen = [self.transformer] + ([cfg_null_model.transformer] if exists(cfg_null_model) else [])
        try:
            self._packed()               # weights are fixed for the whole solve: pack the bf16 operands once (fresh), not 124 times
            if exists(cfg_null_model):
                cfg_null_model._packed()
            for tr in frozen:
                tr.freeze_packed(True)
            fn_eval = fn   # (one function evaluation as a CUDA graph was measured: -1 % at cfg5, but re-capturing per call costs the e2e path 19 %)
            for i in range(steps - 1):
                t0, dt = ts[i], float(ts_host[i + 1] - ts_host[i])
                if method == 'euler':
                    y = ops.axpy(y, fn_eval(t0, y), dt)
                else:
                    half = 0.5 * dt
                    ymid = ops.axpy(y, fn_eval(t0, y), half)
                    y = ops.axpy(y, fn_eval(t0 + half, ymid), dt)
        finally:
            for tr in frozen:
                tr.freeze_packed(False)
        out = torch.where(cond_mask, cond, y)
        if exists(return_raw_output) and return_raw_output:
            return out
        if exists(vocoder):
            return vocoder(out.transpose(1, 2))
        return out

    @_on_module_device
    def forward(self, inp, *, text=None, times=None, lens=None, velocity_consistency_model=None, velocity_consistency_delta=1e-5):
        """Flow-matching training objective, e2_tts.py:1468-1595. Returns E2TTSReturn(loss, cond, pred_flow, pred_data, breakdown)."""
        need_velocity_loss = exists(velocity_consistency_model) and self.velocity_consistency_weight > 0.
        if inp.ndim == 2:
            inp = self.mel_spec(inp).transpose(1, 2)
            assert inp.shape[-1] == self.num_channels
        x1 = inp.to(F32).contiguous()
        B, N, C = x1.shape
        dev = self.device
        if torch.is_grad_enabled():
            ops.zero_pool.begin(dev)
        if isinstance(text, list):
            text = self.tokenizer(text).to(dev)
            assert text.shape[0] == B
        if not exists(lens):
            lens = torch.full((B,), N, device=dev)
        mask = lens_to_mask(lens, length=N)
        # RNG draw order of the reference (Appendix C): frac_lengths, span start, x0, times, text drop
        def span():
            frac = torch.zeros((B,), device=dev).float().uniform_(*self.frac_lengths_mask)
            return mask_from_frac_lengths(lens, frac, N)
        rand_span_mask = _rng.draw('span_mask', span) & mask
        x0 = _rng.draw('x0', lambda: torch.randn_like(x1))
        times = _rng.draw('times', lambda: torch.rand((B,), dtype=x1.dtype, device=dev))
        drop_text_cond = _rng.draw('drop_text_cond', lambda: self.training and pyrandom.random() < self.cond_drop_prob)
        span_u8 = rand_span_mask.to(torch.uint8).contiguous()
        times = times.to(F32).contiguous()
        vel_target = None
        if need_velocity_loss:   # :1556-1576 — the EMA model's prediction at t + delta on the same (x0, x1, cond, text-drop coin), no grad
            vcm = velocity_consistency_model
            with torch.no_grad():
                t_d = times + velocity_consistency_delta
                vpk = vcm._packed()
                A_d, _ = ops.stem_prepare(B, N, C, vpk['Cp'], x1=x1, x0=x0, times=t_d, span=span_u8, concat=velocity_consistency_model.concat_cond)
                y_d, vpk = vcm._embed(A_d, B, N, t_d, mask, text, drop_text_cond, vpk)
                vel_target = ops.PredHead.apply(y_d, vcm.to_pred.weight, vcm.to_pred.bias, vpk['pred'])
        pk = self._packed()
        A, cond = ops.stem_prepare(B, N, C, pk['Cp'], x1=x1, x0=x0, times=times, span=span_u8, want_cond=True, concat=self.concat_cond)
        y, pk = self._embed(A, B, N, times, mask, text, drop_text_cond, pk)
        loss, pred, pred_data, parts = ops.FlowLossHead.apply(y, self.to_pred.weight, self.to_pred.bias, pk['pred'], x1, x0, span_u8, vel_target,
                                                              float(self.velocity_consistency_weight) if need_velocity_loss else 0.0)
        breakdown = LossBreakdown(parts[0], parts[1] if need_velocity_loss else self.zero)
        return E2TTSReturn(loss, cond, pred.view(B, N, C), pred_data.view(B, N, C), breakdown)
