"""Stand-in for einx>=0.3.0 — only the 6 call patterns the reference uses
(e2_tts.py:182,191,220,224,319,326,362,480,1086,1539). Test infrastructure only."""
import torch


def less(pattern, a, b):
    assert pattern.replace(' ', '') == 'n,b->bn', pattern
    return a[None, :] < b[:, None]


def greater_equal(pattern, a, b):
    assert pattern.replace(' ', '') == 'n,b->bn', pattern
    return a[None, :] >= b[:, None]


def where(pattern, cond, a, b):
    p = pattern.replace(' ', '')
    if p == 'bn,bnd,->bnd':
        return torch.where(cond[..., None], a, torch.as_tensor(b, dtype=a.dtype, device=a.device))
    if p == 'bn,bnd,bnd->bnd':
        return torch.where(cond[..., None], a, b)
    raise NotImplementedError(pattern)


def divide(pattern, a, b):
    assert pattern.replace(' ', '') == 'bd,b->bd', pattern
    return a / b[:, None]


def multiply(pattern, a, b):
    p = pattern.replace(' ', '')
    if p == 'i,j->ij':
        return a[:, None] * b[None, :]
    if p == 'bnh,bhnd->bhnd':
        return a.permute(0, 2, 1)[..., None] * b
    raise NotImplementedError(pattern)
