"""Mint the golden vectors under tests/golden/ by running the reference's OWN e2_tts.py (build
container only: needs /root/reference). TEST INFRASTRUCTURE.

    python oracle/make_golden.py

The reference draws its randomness internally (x0, times, span mask); oracle/load_reference.py records
those draws so they can be replayed into the oracle and the CUDA path. dropout=0 because dropout masks
cannot be replayed bit-for-bit across implementations (SURVEY §7 "hard parts").
Fixtures (fp32, torch.save):
  e2tts_d128_L2.pt     E2TTS fwd+bwd: state_dict, mel, text ids, lens, x0, times, span_mask -> loss, pred,
                       cond, parameter grads (full for the text-conditioned case, (norm,sum) per parameter for
                       the text-dropped case)
  sample_d128_L2.pt    E2TTS.sample 4-step midpoint end point (cfg_strength 1 -> cond + null pass + APG)
  duration_d128_L2.pt  DurationPredictor fwd+bwd loss + per-parameter grad (norm,sum), and return_loss=False prediction
  melspec.pt           MelSpec(wave) from torchaudio through the reference module
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.load_reference import load_reference, run_reference_forward  # noqa: E402
from oracle import e2tts_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
TKW = dict(dim=128, depth=2, heads=2)
REF_TKW = dict(dropout=0., max_seq_len=256, **TKW)
TEXT = ['Hello', 'Goodbye']


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    model = ref.E2TTS(transformer=dict(**REF_TKW), use_vocos=False)
    sd = O.randomize_zero_init(model.state_dict(), seed=11)
    model.load_state_dict(sd)
    mel = torch.randn(2, 96, 100)
    lens = torch.tensor([96, 70])
    text_ids = O.list_str_to_tensor(TEXT)
    cases = {}
    for name, drop in (('text', False), ('drop', True)):
        torch.manual_seed(7)
        model.zero_grad()
        out, rec = run_reference_forward(ref, model, mel, TEXT, lens=lens, drop_text_cond=drop)
        out.loss.backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        if drop:  # keep the fixture small: full grads only for the text-conditioned case
            grads = {k: torch.stack((g.norm(), g.sum())) for k, g in grads.items()}
        cases[name] = dict(drop_text_cond=drop, loss=out.loss.detach(), pred=out.pred_flow.detach(),
                           cond=out.cond.detach(), pred_data=out.pred_data.detach(), grads=grads, **rec)
    torch.save(dict(transformer=TKW, max_seq_len=256, state_dict={k: v.clone() for k, v in model.state_dict().items()},
                    mel=mel, text=TEXT, text_ids=text_ids, lens=lens, cases=cases), os.path.join(OUT, 'e2tts_d128_L2.pt'))

    # sample(): record y0 (e2_tts.py:1418)
    model.eval()
    holder = {}

    class Rec:
        def __getattr__(self, n):
            return getattr(torch, n)

        def randn_like(self, *a, **k):
            holder['y0'] = torch.randn_like(*a, **k)
            return holder['y0'].clone()

    torch.manual_seed(3)
    cond = mel[:, :24]
    ref.torch = Rec()
    try:
        smp = model.sample(cond, text=TEXT, duration=64, steps=4, return_raw_output=True)
    finally:
        ref.torch = torch
    torch.save(dict(cond=cond, text_ids=text_ids, duration=64, steps=4, cfg_strength=1.0, y0=holder['y0'], out=smp),
               os.path.join(OUT, 'sample_d128_L2.pt'))

    # DurationPredictor
    torch.manual_seed(1)
    dp = ref.DurationPredictor(transformer=dict(**REF_TKW))
    dsd = O.randomize_zero_init(dp.state_dict(), seed=13)
    dp.load_state_dict(dsd)
    torch.manual_seed(5)
    loss = dp(mel, text=TEXT, lens=lens)
    loss.backward()
    torch.manual_seed(5)
    rand_frac = mel.new_zeros(2).uniform_(0, 1)  # e2_tts.py:1082
    dgrads = {k: torch.stack((p.grad.norm(), p.grad.sum())) for k, p in dp.named_parameters() if p.grad is not None}
    dp.eval()
    with torch.no_grad():
        pred = dp(mel, text=TEXT, lens=lens, return_loss=False)
    torch.save(dict(state_dict={k: v.clone() for k, v in dp.state_dict().items()}, mel=mel, text_ids=text_ids, lens=lens,
                    rand_frac=rand_frac, loss=loss.detach(), grads=dgrads, pred=pred), os.path.join(OUT, 'duration_d128_L2.pt'))

    # MelSpec
    torch.manual_seed(9)
    wave = torch.randn(2, 256 * 24) * 0.3
    torch.save(dict(wave=wave, mel=ref.MelSpec()(wave)), os.path.join(OUT, 'melspec.pt'))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
