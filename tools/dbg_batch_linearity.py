"""debug: single process — gradients of the B=4 batch vs the mean of the two B=2 halves (what 2-rank data parallel must produce)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import e2_tts_pytorch_b200 as pkg
import test_gpu_ddp as T
dev = torch.device('cuda:0')
model = T._make(pkg, dev)
data = T._data(dev)
text = pkg.list_str_to_tensor(['Hello', 'Goodbye', 'Good morning', 'Hi']).to(dev)
g_all, l_all = T._grads_of(pkg, model, slice(0, 4), data, text, False)
g0, l0 = T._grads_of(pkg, model, slice(0, 2), data, text, False)
g1, l1 = T._grads_of(pkg, model, slice(2, 4), data, text, False)
g_all2, _ = T._grads_of(pkg, model, slice(0, 4), data, text, False)
print('loss', l_all, 0.5 * (l0 + l1))
rows = []
for n in g_all:
    w = g_all[n]; a = 0.5 * (g0[n] + g1[n])
    rows.append((T._rel(a, w), T._rel(g_all2[n], w), float(w.norm()), n))
rows.sort(reverse=True)
for r in rows[:12]: print('halves-vs-full %.4g  rerun-vs-full %.4g  norm %.3e  %s' % r)
