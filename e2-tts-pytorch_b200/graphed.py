"""One CUDA graph per training step: forward + backward of an E2TTS model captured once, replayed every step — on one GPU or as one
replica of a data-parallel job (one process per GPU).

The reference's training step (trainer.py:263-270: `loss, cond, pred = self.model(mel_spec, text=text_inputs, lens=mel_lengths)` then
`self.accelerator.backward(loss)`, gradients all-reduced by DDP) is ~800 kernel launches and ~180 autograd nodes here; on a slow host
the Python / driver side of that costs more than the GPU work. Capturing the step removes the host from the critical path. Per step:
  * the new batch is copied into the static input tensors (`mel`, optional `text` / `lens`) — shapes are fixed at capture time;
  * torch's graph-safe CUDA generator advances on every replay (noise x0, flow times, span masks);
  * dropout seeds are kernel ARGUMENTS and therefore frozen in the graph, so every seeded kernel also adds one DEVICE word
    (`seed_dev` in the args structs, include/b200_e2tts.h "dropout seeds") that the graph's first node (`b200_seed_advance`)
    steps in stream order — no host write, so replays may be enqueued back to back without racing on a pinned seed word;
  * with more than one rank (torch.distributed initialised, or `process_group=` given) the graph ends with ONE gather of all
    parameter gradients into a flat fp32 buffer (x 1/world) and the replay is followed by ONE ncclAllReduce of that buffer
    (optim.GradSync): the data-parallel exchange of SURVEY §8e without DDP's bucket hooks, which cannot be captured cheaply and
    whose NCCL kernels would compete with the persistent compute kernels for SMs all through backward.
Gradients are left in `param.grad` exactly as after `loss.backward()` (+ DDP's averaging when world > 1); run the optimiser after
the call (optim.FusedAdoptEMA takes `step.grad_sync.flat` directly). Do not keep an output of an earlier EAGER forward of the same
model alive while constructing this object: its autograd graph pins the parameters' AccumulateGrad nodes to the default (legacy)
stream, which cannot take part in a capture. `cond_drop_prob` must be 0 or 1 while captured (the text-drop coin is a Python-side
branch: it would be frozen either way).
"""
from __future__ import annotations

import torch

from . import lib
from .optim import GradSync, broadcast_module


class _Loss:
    def __init__(self, loss):
        self.loss = loss


class GraphedTrainStep:
    def __init__(self, model, mel, *, text=None, lens=None, warmup=3, process_group=None, flat_grads=None):
        """model: an E2TTS or a DurationPredictor (NOT wrapped in DistributedDataParallel — the exchange is done here). flat_grads: True forces the flat
        gradient buffer even on one rank (for the fused optimiser); default = only when world_size > 1."""
        import torch.distributed as dist
        if hasattr(model, 'module') and not hasattr(model, 'transformer'):
            raise ValueError('GraphedTrainStep: pass the bare E2TTS module, not a DistributedDataParallel wrapper '
                             '(gradients are averaged by one flat all-reduce after the replay)')
        if model.training and 0.0 < float(getattr(model, 'cond_drop_prob', 0.0)) < 1.0:
            raise ValueError('GraphedTrainStep: cond_drop_prob must be 0 or 1 (the text-drop branch is decided on the host)')
        if not mel.is_cuda:
            raise ValueError('GraphedTrainStep: inputs must live on the GPU')
        self.model = model
        dev = mel.device
        self.mel = mel.clone()
        self.text = text.clone() if torch.is_tensor(text) else (model.tokenizer(text).to(dev) if isinstance(text, list) else None)
        self.lens = lens.clone() if torch.is_tensor(lens) else None
        world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            broadcast_module(model, 0, process_group)    # replicas must start identical (DDP does this when it wraps the module)
        use_flat = (world > 1) if flat_grads is None else bool(flat_grads) or world > 1
        self.grad_sync = GradSync(list(model.parameters()), process_group) if use_flat else None
        # the device seed word: every dropout seed of the step is `host seed + *seed_dev` (frozen host part, stepping device part)
        self._seed_dev = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
        model.transformer._seed_dev = self._seed_dev
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up off the capture stream: lazy initialisation, allocator pools, packed weights
            for _ in range(max(1, warmup)):
                self._eager()
                self._clear_grads()
            if self.grad_sync is not None and world > 1:
                self.grad_sync.all_reduce()    # NCCL communicator / channel setup outside the timed path
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()   # the warm-up's side-stream blocks would sit beside the graph's private pool (a full set of activations each)
        self.graph = torch.cuda.CUDAGraph()
        table = self.grad_sync.new_table() if self.grad_sync is not None else None
        n0 = lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = self._eager()
            if self.grad_sync is not None:
                static_grads = self.grad_sync.gather(table)   # recorded against the (still empty) chunk table
        if self.grad_sync is not None:
            self.grad_sync.fill_table(table, static_grads)    # the graph's static gradient tensors, known only now
            self._table = table
        self.launches_per_step = lib.launch_count() - n0   # kernel nodes of ours in the graph (they all run on every replay)
        # the gradient tensors the graph writes into: re-attached after every replay in case the training loop dropped them
        # (optimizer.zero_grad(set_to_none=True)); each replay OVERWRITES them, exactly like backward() into empty .grad fields
        self._grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]
        if self.grad_sync is not None:
            self.grad_sync.attach()

    def _eager(self):
        lib.call('b200_seed_advance', self._seed_dev, torch.cuda.current_stream().cuda_stream)
        out = self.model(self.mel, text=self.text, lens=self.lens)
        if torch.is_tensor(out):       # DurationPredictor.forward returns the scalar loss itself (e2_tts.py:1113)
            out = _Loss(out)
        out.loss.backward()
        return out

    def _clear_grads(self):
        for p in self.model.parameters():
            p.grad = None

    def __call__(self, mel=None, *, text=None, lens=None):
        """Run one step on a new batch of the captured shapes; returns the (device) loss tensor of that step (this rank's)."""
        if mel is not None:
            self.mel.copy_(mel, non_blocking=True)
        if text is not None:
            self.text.copy_(text, non_blocking=True)
        if lens is not None:
            self.lens.copy_(lens, non_blocking=True)
        self.graph.replay()
        if self.grad_sync is not None:
            self.grad_sync.all_reduce()
            self.grad_sync.attach()
        else:
            for p, g in self._grads:
                if p.grad is not g:
                    p.grad = g
        return self.out.loss
