"""Whole-training-step CUDA graph (streams and graphs instead of a tracing compiler).

A MAGVIT2 training step is ~1000 kernel launches, many of them a few microseconds long; launched from
Python the GPU idles between them. Every shape, pointer and hyper-parameter of the step is static, so
forward + backward + fused AdamW are captured ONCE into a CUDA graph and replayed per batch:

    step = GraphedTrainStep(model, optimizer, example_batch)
    loss = step(batch)            # copies the batch into the static input, replays, returns the loss tensor

The step counter AdamW needs for its bias correction lives in device memory (csrc/optim.cu) so the replay
stays exact. Gradient accumulators live in the step-scoped zero arena (ops._ZeroArena: one fill per replay),
activations in the graph's private memory pool; both are reused across replays.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch: torch.Tensor, warmup: int = 3, reducer=None):
        assert example_batch.is_cuda, 'GraphedTrainStep: example batch must be a CUDA tensor'
        from . import ops
        ops.enable_zero_arena(True)   # the captured step owns its gradients: one fill instead of ~540 (ops._ZeroArena)
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.static_in = example_batch.detach().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):          # populate caches, optimizer state, size the zero arena ...
                self._eager_step(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step(self.static_in, zero=False)
        self.replays = 0

    def _eager_step(self, batch, zero=True):
        loss = self.model.training_step(batch, 0)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        return loss.detach()

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        if batch.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(batch, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_loss
