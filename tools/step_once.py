"""Developer tool (GPU): ONE eager cfg2 training step (E2TTS forward + backward, B16 x N1024, d512 / depth 8) inside a cudaProfiler range,
after untimed warm-up steps, so that `ncu --profile-from-start off ...` captures exactly one step's launches.
usage: ncu --profile-from-start off [...] python tools/step_once.py [config 2|3] [dropout]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import e2_tts_pytorch_b200 as pkg

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
dev = torch.device('cuda:0')
torch.manual_seed(0)
dim, depth, heads, B, N = (512, 8, 8, 16, 1024) if cfg == 2 else (1024, 24, 16, 8, 2048)
model = pkg.E2TTS(transformer=dict(dim=dim, depth=depth, heads=heads, dropout=drop), use_vocos=False).to(dev).train()
model.cond_drop_prob = 0.0
mel = torch.randn(B, N, 100, device=dev)
text = pkg.list_str_to_tensor((['Hello', 'Goodbye'] * B)[:B]).to(dev)


def step():
    out = model(mel, text=text)
    out.loss.backward()
    for p in model.parameters():
        p.grad = None


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
