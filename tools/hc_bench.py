"""Developer tool (GPU): device time of the hyper-connection width kernels (unfused path) at the BASELINE shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import e2_tts_pytorch_b200 as pkg
from e2_tts_pytorch_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
for (B, Np, D) in ((16, 1056, 512), (16, 1056, 256), (8, 2080, 1024)):
    S, T = 4, B * Np
    x = (torch.randn(T, S, D, device=dev)).to(torch.bfloat16).requires_grad_()
    P = [torch.randn(D, device=dev) * 0.1, torch.randn(D, S + 1, device=dev) * 0.05, torch.tensor(0.5, device=dev), torch.randn(S, S + 1, device=dev),
         torch.randn(D, device=dev) * 0.05, torch.tensor(0.5, device=dev), torch.randn(S, device=dev)]
    P = [p.requires_grad_() for p in P]
    gain = (1 + 0.1 * torch.randn(B, D, device=dev)).requires_grad_()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tf, tb = [], []
    for it in range(13):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        br, res, beta = ops.HcWidth.apply(x, *P, gain, 2, Np)
        e[1].record()
        gb, gr, gbe = torch.ones_like(br), torch.ones_like(res), torch.ones_like(beta)
        flush.zero_()
        e[2].record()
        torch.autograd.grad([br, res, beta], [x] + P + [gain], [gb, gr, gbe])
        e[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(e[0].elapsed_time(e[1]) * 1e3); tb.append(e[2].elapsed_time(e[3]) * 1e3)
    tf.sort(); tb.sort()
    fb, bb = 9 * D * 2 * T, 13 * D * 2 * T
    print(f'B{B} Np{Np} D{D}: width fwd {tf[len(tf)//2]:.1f} us = {fb / tf[len(tf)//2] * 1e-3:.0f} GB/s, '
          f'bwd (token kernel + param GEMM + finalize + zero slab) {tb[len(tb)//2]:.1f} us = {bb / tb[len(tb)//2] * 1e-3:.0f} GB/s')
