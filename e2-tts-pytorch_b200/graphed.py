"""One CUDA graph per training step: forward + backward of an E2TTS model captured once, replayed every step.

The reference's training step (trainer.py: `loss, cond, pred = self.model(mel_spec, text=text_inputs, lens=mel_lengths)` then
`self.accelerator.backward(loss)`) is ~800 kernel launches and ~180 autograd nodes here; on a slow host the Python / driver side of
that costs more than the 32 ms of GPU work. Capturing the step removes the host from the critical path. What stays per step:
  * the new batch is copied into the static input tensors (`mel`, optional `text` / `lens`) — shapes are fixed at capture time;
  * torch's graph-safe CUDA generator advances on every replay (noise x0, flow times, span masks);
  * dropout seeds are kernel ARGUMENTS and therefore frozen in the graph, so every seeded kernel also adds one device word
    (`b200_set_dropout_seed_device`, include/b200_e2tts.h) that is rewritten from the host RNG before each replay.
Gradients are left in `param.grad` exactly as after `loss.backward()`; run the optimiser after the call. Do not keep an output of an
earlier EAGER forward of the same model alive while constructing this object: its autograd graph pins the parameters' AccumulateGrad
nodes to the default (legacy) stream, which cannot take part in a capture. `cond_drop_prob` must be 0
or 1 while captured (the text-drop coin is a Python-side branch: it would be frozen either way).
"""
from __future__ import annotations

import torch

from . import lib


class GraphedTrainStep:
    def __init__(self, model, mel, *, text=None, lens=None, warmup=3):
        inner = getattr(model, 'module', model)   # a DistributedDataParallel wrapper is called as is; its attributes live on .module
        if inner.training and 0.0 < float(inner.cond_drop_prob) < 1.0:
            raise ValueError('GraphedTrainStep: cond_drop_prob must be 0 or 1 (the text-drop branch is decided on the host)')
        if not mel.is_cuda:
            raise ValueError('GraphedTrainStep: inputs must live on the GPU')
        self.model = model
        dev = mel.device
        self.mel = mel.clone()
        self.text = text.clone() if torch.is_tensor(text) else (inner.tokenizer(text).to(dev) if isinstance(text, list) else None)
        self.lens = lens.clone() if torch.is_tensor(lens) else None
        self._seed_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up off the capture stream: lazy initialisation, allocator pools, packed weights
            for _ in range(max(1, warmup)):
                self._eager()
                self._clear_grads()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        lib.call('b200_set_dropout_seed_device', self._seed_dev)
        n0 = lib.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                self.out = self._eager()
        finally:
            lib.call('b200_set_dropout_seed_device', None)
        self.launches_per_step = lib.launch_count() - n0   # kernel nodes of ours in the graph (they all run on every replay)
        # the gradient tensors the graph writes into: re-attached after every replay in case the training loop dropped them
        # (optimizer.zero_grad(set_to_none=True)); each replay OVERWRITES them, exactly like backward() into empty .grad fields
        self._grads = [(p, p.grad) for p in self.model.parameters() if p.grad is not None]

    def _eager(self):
        out = self.model(self.mel, text=self.text, lens=self.lens)
        out.loss.backward()
        return out

    def _clear_grads(self):
        for p in self.model.parameters():
            p.grad = None

    def __call__(self, mel=None, *, text=None, lens=None):
        """Run one step on a new batch of the captured shapes; returns the (device) loss tensor of that step."""
        if mel is not None:
            self.mel.copy_(mel, non_blocking=True)
        if text is not None:
            self.text.copy_(text, non_blocking=True)
        if lens is not None:
            self.lens.copy_(lens, non_blocking=True)
        self._seed_host.random_(0, 2 ** 62)
        self._seed_dev.copy_(self._seed_host, non_blocking=True)
        self.graph.replay()
        for p, g in self._grads:
            if p.grad is not g:
                p.grad = g
        return self.out.loss
