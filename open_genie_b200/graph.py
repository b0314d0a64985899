"""Whole-training-step CUDA graph (streams and graphs instead of a tracing compiler).

A MAGVIT2 training step is ~860 kernel launches, many of them a few microseconds long; launched from
Python the GPU idles between them. Every shape, pointer and hyper-parameter of the step is static, so
forward + backward (+ the gradient all-reduce for N > 1) + fused AdamW are captured ONCE into a CUDA graph and
replayed per batch:

    step = GraphedTrainStep(model, optimizer, example_batch[, reducer=ArenaGradAllReducer(...)])
    loss = step(batch)            # copies the batch into the static input, replays, returns the loss tensor

What keeps the replay exact and safe:
  * the step counter AdamW needs for its bias correction and the learning rate live in device memory
    (csrc/optim.cu); `optimizer.sync_lr()` refreshes the latter before every replay, so LR schedulers keep working;
  * the step owns a PRIVATE ops.StepScope — its zero arena (gradient accumulators, one fill per replay) and its
    split-K workspace are frozen after capture and stay alive as long as this object does, so later eager calls on
    the same model (validation, tokenize(), a larger batch) can neither free nor overwrite memory whose address is
    baked into the graph, and never allocate from the graph's arena;
  * activations live in the graph's private memory pool;
  * with a reducer, the NCCL all-reduces issued from the autograd hooks are captured too (torch's process group
    forks / joins its communication stream inside the capture), so N > 1 runs the same program as N = 1. The capture
    uses `capture_error_mode='thread_local'`: the NCCL watchdog thread polls events concurrently, which the default
    global mode treats as a capture violation (that is what made the first attempt in round 1 hang).
"""
from __future__ import annotations

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, optimizer, example_batch: torch.Tensor, warmup: int = 3, reducer=None):
        assert example_batch.is_cuda, 'GraphedTrainStep: example batch must be a CUDA tensor'
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.scope = ops.StepScope(ops.ZeroArena())
        if reducer is not None and hasattr(reducer, 'bind_arena'):
            reducer.bind_arena(self.scope.arena)
        self.static_in = example_batch.detach().clone()
        optimizer.zero_grad(set_to_none=True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with ops.step_scope(self.scope):
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 3)):      # populate caches, optimizer state; size + consolidate the arena
                    self._eager_step(self.static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
                self.static_loss = self._eager_step(self.static_in, zero=False)
        self.scope.freeze()
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
        if hasattr(self.optimizer, 'sync_lr'):
            self.optimizer.sync_lr()
        self.graph.replay()
        self.replays += 1
        return self.static_loss
