"""CUDA-graph replay of a whole ELBO-path step.

One training step of this path is ~10 hand-written kernels plus a few dozen tiny torch kernels
(softplus of the kernel hyper-parameters, autograd bookkeeping).  Launched eagerly, the host needs
longer to enqueue them than a B200 needs to run them, so the step is captured once into a CUDA
graph and replayed: every C-ABI call enqueues on torch's current stream, which inside
`torch.cuda.graph` is the capturing stream, so the custom kernels become graph nodes like any
other.  Requirements on `fn`: no host synchronisation (`hlvae_b200.config.check_errors = False`,
subject layouts built beforehand), inputs read from tensors that stay alive ("static" buffers
the caller refills with `copy_` between replays), state carried across steps (m, H, ...) written
back in place.
"""
from __future__ import annotations

import torch


class StepGraph:
    """Capture `fn()` (forward + backward + updates) after `warmup` eager calls; `replay()` runs it.

    `fn` returns a tensor or tuple of tensors; the captured instances are kept in `self.outputs`
    and overwritten by every replay."""

    def __init__(self, fn, warmup: int = 3, stream: torch.cuda.Stream | None = None):
        self.fn = fn
        side = stream if stream is not None else torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = fn()

    def replay(self):
        self.graph.replay()
        return self.outputs
