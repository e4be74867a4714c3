"""CPU: the oracle restatement (oracle/e2tts_oracle.py) against the golden vectors minted from the
reference's own e2_tts.py (oracle/make_golden.py). Tolerance: fp32 rounding (rel-L2 <= 1e-4)."""
import os

import pytest
import torch

from oracle import e2tts_oracle as O
from conftest import rel_l2, GOLDEN


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.mark.parametrize('case', ['text', 'drop'])
def test_e2tts_forward_backward_matches_golden(case):
    g = _load('e2tts_d128_L2.pt')
    c = g['cases'][case]
    cfg = O.TransformerCfg(**g['transformer'])
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g['state_dict'].items()}
    out = O.e2tts_forward(sd, cfg, g['mel'], g['text_ids'], x0=c['x0'], times=c['times'], span_mask=c['span_mask'],
                          lens=g['lens'], drop_text_cond=c['drop_text_cond'])
    assert rel_l2(out['pred'], c['pred']) < 1e-4
    assert rel_l2(out['cond'], c['cond']) == 0.0
    assert abs(out['loss'].item() - c['loss'].item()) < 1e-4 * abs(c['loss'].item())
    out['loss'].backward()
    for k, gref in c['grads'].items():
        got = sd[k].grad
        assert got is not None, k
        if case == 'drop':
            got = torch.stack((got.norm(), got.sum()))
        assert (got - gref).abs().max() <= 2e-4 * gref.abs().max() + 1e-7, k
    if case == 'drop':  # text-stream parameters must receive no gradient (DDP find_unused semantics)
        for k, v in sd.items():
            is_text = ('text' in k) or ('.layers.' in k and k.split('.layers.')[1].split('.')[1] == '1') \
                or ('.hyper_conns.' in k and k.split('.hyper_conns.')[1].split('.')[1] == '1')
            if is_text and v.requires_grad:
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k


def test_sample_matches_golden():
    g = _load('sample_d128_L2.pt')
    e = _load('e2tts_d128_L2.pt')
    cfg = O.TransformerCfg(**e['transformer'])
    out = O.e2tts_sample(e['state_dict'], cfg, g['cond'], g['text_ids'], duration=g['duration'], y0=g['y0'],
                         steps=g['steps'], cfg_strength=g['cfg_strength'])
    assert out.shape == g['out'].shape
    assert rel_l2(out, g['out']) < 1e-4


def test_duration_matches_golden():
    g = _load('duration_d128_L2.pt')
    e = _load('e2tts_d128_L2.pt')
    cfg = O.TransformerCfg(cond_on_time=False, **e['transformer'])
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g['state_dict'].items()}
    loss = O.duration_forward(sd, cfg, g['mel'], g['text_ids'], lens=g['lens'], rand_frac=g['rand_frac'])
    assert abs(loss.item() - g['loss'].item()) < 1e-4 * abs(g['loss'].item())
    loss.backward()
    total = torch.cat([v.flatten() for v in g['grads'].values()]).norm()
    for k, gref in g['grads'].items():   # full per-parameter gradients of the reference (fixture re-minted in round 2, oracle/make_golden.py)
        assert (sd[k].grad - gref).norm() <= 5e-4 * gref.norm() + 1e-6 * total, k
    with torch.no_grad():
        pred = O.duration_forward(g['state_dict'], cfg, g['mel'], g['text_ids'], lens=g['lens'], return_loss=False)
    assert rel_l2(pred, g['pred']) < 1e-5


def test_melspec_matches_golden():
    g = _load('melspec.pt')
    out = O.melspec(g['wave'])
    assert out.shape == g['mel'].shape
    assert (out - g['mel']).abs().max() < 1e-3


def test_tokenizer_and_masks():
    ids = O.list_str_to_tensor(['Hello', 'Goodbye'])
    assert ids.tolist() == [[72, 101, 108, 108, 111, -1, -1], [71, 111, 111, 100, 98, 121, 101]]
    m = O.lens_to_mask(torch.tensor([3, 1]), 4)
    assert m.tolist() == [[True, True, True, False], [True, False, False, False]]
