"""Developer tool (GPU): time the short-K / epilogue-heavy GEMM shapes of cfg2 with individual epilogue features switched off, to see
what the epilogue costs (CUDA events, L2 flushed between launches by cycling through several operand sets)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import e2_tts_pytorch_b200 as pkg
from e2_tts_pytorch_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
NSET = 6


def bench(name, M, N, K, reps=30, **kw):
    geglu = kw.get('geglu', False)
    sets = []
    for _ in range(NSET):
        A, Bm = bf(M, K), bf(N, K)
        out = torch.empty((M, N // 2 if geglu else N), device=dev, dtype=torch.bfloat16)
        d = dict(out=out)
        if kw.get('d2'):
            d['D2'] = torch.empty((M, N), device=dev, dtype=torch.bfloat16); d['ldd2'] = N
        if kw.get('bias'):
            d['bias'] = torch.randn(N, device=dev)
        if kw.get('colscale'):
            d['colscale'] = torch.rand(16, N, device=dev); d['rows_per_batch'] = M // 16
        if kw.get('rowmask'):
            d['rowmask'] = torch.ones(M, dtype=torch.uint8, device=dev)
        if kw.get('resid'):
            d['resid'] = bf(M, N); d['ldr'] = N
        if geglu:
            d['geglu'] = True; d['dropout_p'] = kw.get('p', 0.0); d['seed'] = 5
        if kw.get('tile'):
            d['force_tile'] = kw['tile']
        sets.append((A, Bm, d))
    def run(i):
        A, Bm, d = sets[i % NSET]
        ops.gemm(A, Bm, M, N, K, **d)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(reps))
    med = ts[len(ts) // 2]
    print(f'{name:58s} {M}x{N}x{K}: {med:7.1f} us  {2.0 * M * N * K / med / 1e6:7.1f} TF/s', flush=True)


T = 16896
bench('FF-in GEGLU bias drop0.1 D2 (as in the step)', T, 4096, 512, geglu=True, bias=True, p=0.1, d2=True)
bench('FF-in GEGLU bias drop0 D2', T, 4096, 512, geglu=True, bias=True, p=0.0, d2=True)
bench('FF-in GEGLU bias drop0 no-D2', T, 4096, 512, geglu=True, bias=True, p=0.0)
bench('FF-in GEGLU nobias drop0 no-D2', T, 4096, 512, geglu=True)
bench('FF-in plain N=4096 (no GEGLU)', T, 4096, 512)
bench('text FF-in GEGLU bias drop0.1 D2', T, 2048, 256, geglu=True, bias=True, p=0.1, d2=True)
bench('text FF-in GEGLU nobias drop0 no-D2', T, 2048, 256, geglu=True)
bench('text FF-in plain N=2048', T, 2048, 256)
bench('out-proj colscale+rowmask', T, 512, 512, colscale=True, rowmask=True)
bench('out-proj plain', T, 512, 512)
bench('out-proj plain tile 256x128', T, 512, 512, tile=2)
bench('out-proj plain tile 128x128', T, 512, 512, tile=1)
bench('text out-proj rowmask (K=512,N=256)', T, 256, 512, rowmask=True)
bench('text out-proj tile 256x128', T, 256, 512, rowmask=True, tile=2)
bench('plain K=256 N=512', T, 512, 256)
bench('plain K=256 N=512 tile 256x128', T, 512, 256, tile=2)
bench('plain K=256 N=512 tile 128x128', T, 512, 256, tile=1)
bench('qkv N=1552 K=512', T, 1552, 512)
bench('qkv text N=1552 K=256', T, 1552, 256)
bench('FF-out K=2048 bias colscale', T, 512, 2048, bias=True, colscale=True)
bench('cross two-src-like K=768 resid (4T rows)', 4 * T, 512, 768, resid=True)
