"""CPU, build container only: pins oracle/e2tts_oracle.py against the reference's own e2_tts.py loaded
unmodified (oracle/load_reference.py). Skipped where /root/reference does not exist (the GPU box)."""
import pytest
import torch

from oracle import e2tts_oracle as O
from oracle.load_reference import load_reference, reference_available, run_reference_forward
from conftest import rel_l2

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')


@pytest.mark.parametrize('depth,lens', [(2, None), (4, [80, 51])])
def test_forward_backward_vs_reference(depth, lens):
    ref = load_reference()
    torch.manual_seed(depth)
    kw = dict(dim=128, depth=depth, heads=2)
    model = ref.E2TTS(transformer=dict(dropout=0., max_seq_len=128, **kw), use_vocos=False)
    model.load_state_dict(O.randomize_zero_init(model.state_dict(), seed=depth))
    mel = torch.randn(2, 80, 100)
    lens_t = torch.tensor(lens) if lens else None
    text = ['abc', 'a longer text than the first']
    out, rec = run_reference_forward(ref, model, mel, text, lens=lens_t)
    out.loss.backward()
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    o = O.e2tts_forward(sd, O.TransformerCfg(**kw), mel, O.list_str_to_tensor(text), lens=lens_t, **rec)
    o['loss'].backward()
    assert rel_l2(o['pred'], out.pred_flow) < 1e-4
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert (p.grad - sd[k].grad).abs().max() <= 2e-4 * p.grad.abs().max() + 1e-7, k


def test_melspec_vs_torchaudio():
    ref = load_reference()
    wave = torch.randn(1, 256 * 10 + 17)
    assert (ref.MelSpec()(wave) - O.melspec(wave)).abs().max() < 1e-3
