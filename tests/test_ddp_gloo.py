"""CPU, world_size 2, gloo: host-side logic of the N > 1 path. The CUDA kernels cannot run here, so the
module's parameters are driven by the fp32 oracle (same state_dict) under torch DDP: this checks that (1) the parameter
tree of our E2TTS is DDP-wrappable and receives per-parameter gradients, (2) rank-sharded batches + gradient averaging equal
the single-process average of the two per-rank losses (the reference's per-rank masked-mean semantics, e2_tts.py:1582 /
trainer.py:270), and (3) dropping the text on a step leaves text-stream parameters without gradient
(find_unused_parameters=True semantics, trainer.py:155)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, drop, q):
    try:
        _worker_impl(rank, world, port, drop, q)
    except Exception as e:  # surface failures immediately instead of a queue timeout
        import traceback
        q.put(('error', traceback.format_exc()))
        raise


def _worker_impl(rank, world, port, drop, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    import e2_tts_pytorch_b200 as pkg
    from oracle import e2tts_oracle as O
    torch.manual_seed(0)
    model = pkg.E2TTS(transformer=dict(dim=128, depth=2, heads=1, dropout=0., max_seq_len=64), use_vocos=False)
    with torch.no_grad():
        O.randomize_zero_init(dict(model.named_parameters()), seed=3)
    cfg = O.TransformerCfg(dim=128, depth=2, heads=1)

    class OracleDriven(torch.nn.Module):   # same parameters, forward through the oracle (CPU stand-in for the CUDA path)
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, mel, text, x0, times, span):
            sd = dict(self.m.named_parameters())
            sd.update(dict(self.m.named_buffers()))
            return O.e2tts_forward(sd, cfg, mel, text, x0=x0, times=times, span_mask=span, drop_text_cond=drop)['loss']

    ddp = torch.nn.parallel.DistributedDataParallel(OracleDriven(model), find_unused_parameters=True)
    g = torch.Generator().manual_seed(123)
    mel = torch.randn(4, 48, 100, generator=g)
    x0 = torch.randn(4, 48, 100, generator=g)
    times = torch.rand(4, generator=g)
    span = torch.zeros(4, 48, dtype=torch.bool)
    span[:, 5:40] = True
    text = O.list_str_to_tensor(['ab', 'cde', 'f', 'ghij'])
    sl = slice(rank * 2, rank * 2 + 2)
    loss = ddp(mel[sl], text[sl], x0[sl], times[sl], span[sl])
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    if rank == 0:
        # single-process reference: mean of the two per-rank losses
        for p in model.parameters():
            p.grad = None
        plain = OracleDriven(model)
        tot = 0.5 * (plain(mel[:2], text[:2], x0[:2], times[:2], span[:2]) + plain(mel[2:], text[2:], x0[2:], times[2:], span[2:]))
        tot.backward()
        worst, unused_ok = 0.0, True
        for k, p in model.named_parameters():
            if p.grad is None:
                unused_ok &= grads[k] is None or float(grads[k].abs().max()) == 0.0
                continue
            worst = max(worst, float((grads[k] - p.grad).abs().max() / (p.grad.abs().max() + 1e-12)))
        text_unused = all(grads[k] is None or float(grads[k].abs().max()) == 0.0
                          for k in grads if k.startswith('embed_text') or 'text_registers' in k)
        q.put((worst, unused_ok, text_unused))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('drop', [False, True])
def test_ddp_gradient_averaging_and_unused_text_params(drop):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, drop, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    assert res[0] != 'error', res[1]
    worst, unused_ok, text_unused = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst < 1e-4, worst
    assert unused_ok
    assert text_unused == drop
