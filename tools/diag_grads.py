"""Developer diagnostic (GPU): per-parameter gradient norm ratio / cosine of the CUDA path vs the golden vectors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import e2_tts_pytorch_b200 as pkg
from oracle import e2tts_oracle as O

dev = torch.device('cuda:0')
G = os.path.join(ROOT, 'tests', 'golden')
e = torch.load(os.path.join(G, 'e2tts_d128_L2.pt'), weights_only=False)
c = e['cases']['text']
m = pkg.E2TTS(transformer=dict(dropout=0., max_seq_len=e['max_seq_len'], **e['transformer']), use_vocos=False)
m.load_state_dict(e['state_dict']); m.to(dev).train()
with pkg.inject_randomness(x0=c['x0'].to(dev), times=c['times'].to(dev), span_mask=c['span_mask'].to(dev), drop_text_cond=False):
    out = m(e['mel'].to(dev), text=e['text'], lens=e['lens'].to(dev))
out.loss.backward()
print('e2tts loss', float(out.loss), float(c['loss']))
rows = []
for k, p in m.named_parameters():
    if k in c['grads']:
        g, r = p.grad.cpu().double().flatten(), c['grads'][k].double().flatten()
        rows.append((float(g.norm() / (r.norm() + 1e-30)), float(g @ r / (g.norm() * r.norm() + 1e-30)), float(r.norm()), k))
rows.sort()
for r in rows[:12] + rows[-8:]:
    print('  ratio %.3f cos %.4f refnorm %.3e %s' % r)

d = torch.load(os.path.join(G, 'duration_d128_L2.pt'), weights_only=False)
dp = pkg.DurationPredictor(transformer=dict(dropout=0., max_seq_len=256, **e['transformer']))
dp.load_state_dict(d['state_dict']); dp.to(dev).train()
with pkg.inject_randomness(duration_rand_frac=d['rand_frac'].to(dev)):
    loss = dp(d['mel'].to(dev), text=e['text'], lens=d['lens'].to(dev))
loss.backward()
print('duration loss', float(loss), float(d['loss']))
# oracle grads for cosine
sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in d['state_dict'].items()}
lo = O.duration_forward(sd, O.TransformerCfg(cond_on_time=False, **e['transformer']), d['mel'], d['text_ids'], lens=d['lens'], rand_frac=d['rand_frac'])
lo.backward()
rows = []
for k, p in dp.named_parameters():
    if sd[k].grad is not None and p.grad is not None:
        g, r = p.grad.cpu().double().flatten(), sd[k].grad.double().flatten()
        rows.append((float(g.norm() / (r.norm() + 1e-30)), float(g @ r / (g.norm() * r.norm() + 1e-30)), float(r.norm()), k))
rows.sort()
for r in rows[:15] + rows[-10:]:
    print('  ratio %.3f cos %.4f refnorm %.3e %s' % r)
