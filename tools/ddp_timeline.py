"""Developer tool (2+ GPUs, torchrun): where the data-parallel step's time goes. Per rank, CUDA events bracket (a) the replay of the
captured forward + backward + flat gradient gather and (b) the ONE ncclAllReduce of the flat gradient buffer that follows it
(e2_tts_pytorch_b200.GraphedTrainStep / GradSync), over `iters` steps after warm-up; rank 0 prints per-rank medians and the exposed
share of the exchange. The all-reduce time of a rank includes waiting for the slowest rank's replay (skew).
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_timeline.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import e2_tts_pytorch_b200 as pkg

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
model = pkg.E2TTS(transformer=dict(dim=512, depth=8, heads=8, dropout=0.1), use_vocos=False).to(dev).train()
model.cond_drop_prob = 0.0
torch.manual_seed(rank)
mel = torch.randn(16, 1024, 100, device=dev)
text = pkg.list_str_to_tensor((['Hello', 'Goodbye'] * 8)).to(dev)
step = pkg.GraphedTrainStep(model, mel, text=text)
for _ in range(5):
    step()
torch.cuda.synchronize()
dist.barrier()
rep, ar, tot = [], [], []
for _ in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    step.graph.replay()
    e[1].record()
    step.grad_sync.all_reduce()
    e[2].record()
    torch.cuda.synchronize()
    rep.append(e[0].elapsed_time(e[1])); ar.append(e[1].elapsed_time(e[2])); tot.append(e[0].elapsed_time(e[2]))
med = lambda v: sorted(v)[len(v) // 2]
mine = torch.tensor([med(rep), med(ar), med(tot)], device=dev)
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
if rank == 0:
    nbytes = step.grad_sync.flat.numel() * 4
    print(f'data-parallel step timeline, {world} ranks, cfg2 per rank (B16 x N1024), flat gradient buffer {nbytes / 1e6:.1f} MB fp32')
    for r, v in enumerate(allv):
        a, b, c = (float(x) for x in v)
        print(f'  rank {r}: graph replay (fwd + bwd + gather) {a:7.3f} ms | ncclAllReduce {b:6.3f} ms ({nbytes / b / 1e6:7.1f} GB/s algorithmic) | step {c:7.3f} ms | exchange share {100 * b / c:4.1f} %')
dist.destroy_process_group()
