"""Developer tool (GPU): run the non-GEMM hot kernels once at cfg2 shapes so `ncu --set full -k regex:...` captures are short."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import e2_tts_pytorch_b200 as pkg
from e2_tts_pytorch_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
B, Np, S, D, H = 16, 1056, 4, 512, 8
T = B * Np
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
which = sys.argv[1] if len(sys.argv) > 1 else 'hc,conv,attn'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if 'hc' in which:
    x = bf(T, S, D).requires_grad_()
    P = [torch.randn(D, device=dev) * 0.1, torch.randn(D, S + 1, device=dev) * 0.05, torch.tensor(0.5, device=dev), torch.randn(S, S + 1, device=dev),
         torch.randn(D, device=dev) * 0.05, torch.tensor(0.5, device=dev), torch.randn(S, device=dev)]
    P = [p.requires_grad_() for p in P]
    gain = (1 + 0.1 * torch.randn(B, D, device=dev)).requires_grad_()
    y = bf(T, D)
    for _ in range(reps):
        br, res, beta = ops.HcWidth.apply(x, *P, gain, 2, Np)
        out = ops.HcDepth.apply(res, y, beta)
        torch.autograd.grad([br, out], [x], [torch.ones_like(br), torch.ones_like(out)])
if 'conv' in which:
    x = bf(T, D).requires_grad_()
    w = (torch.randn(D, 1, 31, device=dev) * 0.1).requires_grad_()
    b = torch.zeros(D, device=dev).requires_grad_()
    m = torch.ones(B, Np, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        yy = ops.DwConv.apply(x, w, b, m, B, Np)
        torch.autograd.grad(yy, [x, w, b], torch.ones_like(yy))
if 'attn' in which:
    q, k, v = (bf(B, H, Np, 64).requires_grad_() for _ in range(3))
    gate = torch.rand(T, H, device=dev).requires_grad_()
    m = torch.ones(B, Np, dtype=torch.uint8, device=dev)
    for _ in range(reps):
        og = ops.AttnCore.apply(q, k, v, gate, m, 0.1, 7, 50.0, None)
        torch.autograd.grad(og, [q, k, v], torch.ones_like(og))
if 'gemm' in which:   # the epilogue-heavy / short-K GEMMs of the step: FF-in GEGLU (audio, text), out-proj, a plain K=256 problem
    for (N, K, kw) in [(4096, 512, dict(geglu=True, bias=torch.randn(4096, device=dev), dropout_p=0.1, seed=3, D2=torch.empty((T, 4096), device=dev, dtype=torch.bfloat16), ldd2=4096)),
                       (2048, 256, dict(geglu=True, bias=torch.randn(2048, device=dev), dropout_p=0.1, seed=3, D2=torch.empty((T, 2048), device=dev, dtype=torch.bfloat16), ldd2=2048)),
                       (512, 512, dict(colscale=torch.rand(B, 512, device=dev), rows_per_batch=Np, rowmask=torch.ones(T, dtype=torch.uint8, device=dev))),
                       (512, 256, dict())]:
        A, W = bf(T, K), bf(N, K)
        for _ in range(reps):
            ops.gemm(A, W, T, N, K, **kw)
torch.cuda.synchronize()
print('done')
