#!/usr/bin/env python
"""bench.py — mel-frames/sec of the E2TTS flow-matching training step (forward + loss.backward()) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3] [--dropout P]

Workload (BASELINE.json configs[1], "cfg2"): E2TTS(dim 512, depth 8, heads 8, 100 mels), bf16 tensor-core compute,
per-GPU batch 16 x 1024 mel frames, text conditioning on every step (the expensive branch), synthetic data,
random-init weights. One process per GPU (torchrun), DDP gradient all-reduce over NCCL, weak scaling.
One JSON line is printed by rank 0 (contract: see DESIGN.md §measurement):
  value     whole-job mel-frames/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e       same metric through the public API with HOST (pinned) inputs: H2D of mel every step + D2H of the loss
  roofline  tcgen05 GEMM kernel: algorithmic FLOPs / CUDA-event time vs the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port (oracle/e2tts_oracle.py = the reference algorithm in fp32 PyTorch) on the host cores
`--impl reference` times that CPU path alone (the reference itself is pure Python + unvendored deps and cannot travel
to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'mel-frames/sec E2TTS fwd+bwd (d512 depth8 L1024)'
TEXT = ['Hello', 'Goodbye']
CONFIGS = {
    2: dict(dim=512, depth=8, heads=8, batch=16, seq=1024, name='cfg2: E2TTS d512 depth8 h8, B16 x N1024 x 100 mel, fwd+bwd'),
    3: dict(dim=1024, depth=24, heads=16, batch=8, seq=2048, name='cfg3: E2TTS d1024 depth24 h16, B8 x N2048 per GPU, fwd+bwd'),
}


def step_flops(cfg, B):
    """Algorithmic FLOPs of one forward (SURVEY §8d formulas); fwd+bwd = 3x."""
    d, L, h, N = cfg['dim'], cfg['depth'], cfg['heads'], cfg['seq']
    dt, I, S, Np = d // 2, h * 64, 4, N + 32
    T = B * Np
    f = 0.0
    for i in range(L):
        f += 6 * T * d * I + 2 * T * I * d + 4 * T * Np * I + 24 * T * d * d + 2 * T * 31 * d + 2 * T * d * h * (1 if i == 0 else 2)
        if i >= L // 2:
            f += 4 * S * T * d * d
        f += 6 * T * dt * I + 2 * T * I * dt + 4 * T * Np * I + 24 * T * dt * dt + 2 * T * 31 * dt + 2 * T * dt * h * (1 if i == 0 else 2)
        f += 2 * S * T * (d + dt) * d + (2 * S * T * (d + dt) * dt if i != L - 1 else 0)
        f += 2 * 3 * T * (4 * S * d * (S + 1) + 4 * S * d) * 0.75
    f += 4 * B * N * 100 * d + 2 * B * N * d * 100
    return f


# ----------------------------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, batch, threads):
    """The reference algorithm on the host: fp32 oracle port, forward + backward on a `batch`-sample slice."""
    import e2_tts_pytorch_b200 as pkg
    from oracle import e2tts_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = pkg.E2TTS(transformer=dict(dim=cfg['dim'], depth=cfg['depth'], heads=cfg['heads'], dropout=0.), use_vocos=False)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    ocfg = O.TransformerCfg(dim=cfg['dim'], depth=cfg['depth'], heads=cfg['heads'])
    N = cfg['seq']
    mel = torch.randn(batch, N, 100)
    text = O.list_str_to_tensor([TEXT[i % 2] for i in range(batch)])
    span = torch.zeros(batch, N, dtype=torch.bool)
    span[:, N // 10: N - N // 10] = True

    def step():
        out = O.e2tts_forward(sd, ocfg, mel, text, x0=torch.randn_like(mel), times=torch.rand(batch), span_mask=span)
        out['loss'].backward()
        for v in sd.values():
            v.grad = None
        return float(out['loss'])

    return step


def run_cpu(cfg, steps, warmup, batch=1, budget_s=60.0, threads=None):
    threads = threads or min(os.cpu_count() or 1, 32)   # more threads than this only oversubscribes the small per-layer GEMMs
    step = cpu_step_fn(cfg, batch, threads)
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        step()
        dt = time.time() - t0
        if i >= warmup or (time.time() - t_begin > budget_s):
            times.append(dt)
        if time.time() - t_begin > budget_s and times:
            break
    times.sort()
    med = times[len(times) // 2]
    return dict(value=batch * cfg['seq'] / med, unit='mel-frames/s', cores=threads, kind='port',
                sample=f'{batch} of {cfg["batch"]} sequences x {cfg["seq"]} frames per step (same model/seq_len), fp32, '
                       f'median of {len(times)} timed step(s), {threads} host threads', ms_per_step=med * 1e3)


def run_cpu_bounded(config, steps, warmup, timeout_s):
    """Run the CPU leg in a child process with a hard wall-clock bound, so that a slow host can never cost the GPU line."""
    cmd = [sys.executable, os.path.abspath(__file__), '--cpu-worker', '--config', str(config), '--steps', str(steps), '--warmup', str(warmup)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        note = 'cpu worker produced no result: ' + out.stderr.strip()[-200:]
    except subprocess.TimeoutExpired:
        note = f'one fp32 CPU step of the oracle port did not finish within the {timeout_s:.0f} s bound on this host'
    return dict(value=None, unit='mel-frames/s', cores=min(os.cpu_count() or 1, 32), kind='port', sample=note)


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, idx):
        self.idx, self.samples, self.stop = idx, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', f'--id={self.idx}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                    reasons=sorted(reasons), samples=len(self.samples))


def run_gpu(args):
    import torch.distributed as dist
    import e2_tts_pytorch_b200 as pkg
    from e2_tts_pytorch_b200 import lib, ops
    cfg = CONFIGS[args.config]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    graph_ddp = world > 1 and args.graph_ddp
    if graph_ddp:   # EXPERIMENTAL (off by default, not yet validated on hardware): capture the DDP step, NCCL all-reduce included
        os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')   # the watchdog's event queries would invalidate the capture
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = pkg.E2TTS(transformer=dict(dim=cfg['dim'], depth=cfg['depth'], heads=cfg['heads'], dropout=args.dropout), use_vocos=False).to(dev)
    model.train()
    model.cond_drop_prob = 0.0  # text conditioning on every step: the expensive branch, identical graph on every rank
    net = model
    if world > 1:
        if graph_ddp:   # PyTorch's recipe for whole-backward capture under DDP: construct DDP on a side stream
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
            torch.cuda.current_stream(dev).wait_stream(side)
        else:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    B, N = cfg['batch'], cfg['seq']
    torch.manual_seed(rank)
    host_mel = torch.randn(B, N, 100).pin_memory()
    dev_mel = host_mel.to(dev)
    text = [TEXT[i % 2] for i in range(B)]
    text_dev = pkg.list_str_to_tensor(text).to(dev)

    def step(mel, readback):
        out = net(mel, text=text_dev)
        out.loss.backward()
        for p in model.parameters():
            p.grad = None
        return out.loss.item() if readback else None

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(args.warmup, 3)):
        step(dev_mel, False)
    # -- per-kernel-family CUDA-event timing of the tcgen05 GEMM inside real steps (roofline numerator/denominator)
    prof = dict(flops=0.0, events=[])
    orig_gemm = ops.gemm

    def gemm_timed(A, Bm, M, Nn, K, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_gemm(A, Bm, M, Nn, K, **kw)
        e1.record()
        prof['events'].append((e0, e1, (M, Nn, K, int(kw.get('a_mn', False)), int(kw.get('b_mn', False)), int(kw.get('split_k', 1)),
                                        int(bool(kw.get('geglu'))), int(kw.get('A2') is not None))))
        prof['flops'] += 2.0 * M * Nn * K
        return out

    step_mode, graph_note, eager_ms = 'eager', None, None
    with ClockSampler(local) as clk:
        n0 = lib.launch_count()
        ms_dev = timed(lambda: step(dev_mel, False), args.steps)
        launches = (lib.launch_count() - n0) // args.steps
        ms_e2e = timed(lambda: step(host_mel.to(dev, non_blocking=True), True), args.steps)
        ops.gemm = gemm_timed
        nprof = min(args.steps, 3)
        for _ in range(nprof):
            step(dev_mel, False)
        torch.cuda.synchronize()
        ops.gemm = orig_gemm
        # -- the same step through pkg.GraphedTrainStep (forward + backward captured in one CUDA graph): identical kernels and work,
        #    no per-launch host cost. Single-GPU only (DDP's bucketed all-reduce is not captured); falls back to the eager numbers.
        if (world == 1 and not args.no_graph) or graph_ddp:
            try:
                eager_loss = step(dev_mel, True)
                graphed = pkg.GraphedTrainStep(net, dev_mel, text=text_dev, warmup=11 if world > 1 else 3)
                g_loss = float(graphed().item())
                if not (g_loss == g_loss and 0.5 * eager_loss <= g_loss <= 2.0 * eager_loss):
                    raise RuntimeError(f'graphed loss {g_loss} vs eager {eager_loss}')
                for _ in range(3):
                    graphed()
                ms_g = timed(lambda: graphed(), args.steps)
                ms_g_e2e = timed(lambda: graphed(host_mel).item(), args.steps)
                if ms_g < ms_dev:
                    eager_ms, step_mode = ms_dev, 'cuda_graph'
                    ms_dev, ms_e2e, launches = ms_g, ms_g_e2e, graphed.launches_per_step
                else:
                    graph_note = f'captured but not faster ({ms_g:.2f} ms)'
            except Exception as e:  # noqa: BLE001 - any capture problem: keep the eager measurement
                graph_note = f'unavailable: {type(e).__name__}: {str(e)[:120]}'
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in prof['events'])
    if os.environ.get('B200_GEMM_BREAKDOWN') and rank == 0:
        agg = {}
        for a, b, key in prof['events']:
            t = agg.setdefault(key, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(b)
        print('GEMM breakdown over %d profiled steps: (M, N, K, a_mn, b_mn, split, geglu, two_src) count ms TF/s' % nprof, file=sys.stderr)
        for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print('  %-44s n=%4d  %8.3f ms  %7.1f TF/s' % (str(key), n, ms, 2.0 * key[0] * key[1] * key[2] * n / ms * 1e-9), file=sys.stderr)
    n_gemm = len(prof['events'])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    achieved_tf = prof['flops'] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    frames = world * B * N
    fl = 3 * step_flops(cfg, B)
    line = {
        'metric': METRIC, 'value': frames / (ms_dev * 1e-3), 'unit': 'mel-frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'per_gpu_batch': B, 'seq_len': N, 'global_batch': world * B, 'parallelism': f'dp{world}',
                   'dropout': args.dropout, 'text_cond': 'on every step', 'weights': 'random init', 'optimizer_step': 'not part of the metric (fwd+bwd)',
                   'l2': 'per-step working set (~10 GB of activations) >> 126 MB L2, no flush needed',
                   'step': ('E2TTS forward + loss.backward() replayed through e2_tts_pytorch_b200.GraphedTrainStep (one CUDA graph, same kernels)'
                            if step_mode == 'cuda_graph' else 'E2TTS forward + loss.backward(), eager launches'),
                   **({'eager_ms_per_step': eager_ms} if eager_ms is not None else {}), **({'cuda_graph': graph_note} if graph_note else {})},
        'e2e': {'value': frames / (ms_e2e * 1e-3), 'unit': 'mel-frames/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': host_mel.numel() * 4, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'clocks': clk.summary(),
        'roofline': {'kernel': 'gemm_tcgen05_kernel (all GEMMs of the step: fwd, dX, dW)', 'bound': 'tensor', 'achieved': achieved_tf, 'peak': peak_tf,
                     'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf if peak_tf else None, 'traffic': None,
                     'launches_per_step': n_gemm // max(nprof, 1), 'ms_per_step': gemm_ms / max(nprof, 1),
                     'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PF/s sustained',
                     'step_flops': fl, 'step_tensor_frac': fl / (ms_dev * 1e-3) / 1e12 / peak_tf},
    }
    if world == 1 and not args.no_cpu:
        line['cpu_baseline'] = run_cpu_bounded(args.config, steps=1, warmup=1, timeout_s=100.0)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=[2, 3])
    ap.add_argument('--dropout', type=float, default=0.1)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time the eager step only (skip the CUDA-graph replay of the same step)')
    ap.add_argument('--graph-ddp', action='store_true', help='EXPERIMENTAL: also capture the N > 1 DDP step (NCCL all-reduce inside the graph)')
    ap.add_argument('--cpu-worker', action='store_true', help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        print(json.dumps(run_cpu(CONFIGS[args.config], steps=args.steps, warmup=args.warmup, batch=1, budget_s=60.0)), flush=True)
        return
    if args.impl == 'reference':
        if int(os.environ.get('RANK', '0')) != 0:
            return
        cfg = CONFIGS[args.config]
        r = run_cpu_bounded(args.config, steps=max(1, min(args.steps, 2)), warmup=min(args.warmup, 1), timeout_s=170.0)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'mel-frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': r.get('ms_per_step'), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': cfg['name'], 'note': 'reference algorithm on host CPU cores (oracle port)'},
            'cpu_baseline': {k: r.get(k) for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': r['value'], 'unit': 'mel-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    run_gpu(args)


if __name__ == '__main__':
    main()
