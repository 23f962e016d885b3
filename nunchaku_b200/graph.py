"""CUDA-graph capture of a step built from this package's operators.

The reference's forward is not capture-safe (host-side ``zero_()`` of ``lora_act_out`` / ``out_vk`` and ``cudaMallocAsync`` temporaries inside
the ops: gemm_w4a4_launch_impl.cuh:252,487, src/Tensor.h:84-107; a FLUX step is ~530 GEMM-side launches + ~600 glue launches).  Every
operator here is: no host synchronisation, no host-side memsets, split-K workspaces and tickets are self-cleaning, converted weights are built
before the first call, temporaries come from torch's caching allocator (which hands a capture its own pool).  ``GraphedStep`` packages the
usual recipe -- warm up on a side stream, capture once, replay with static input / output buffers:

    step = GraphedStep(lambda x: block(x), (example_x,))
    y = step(x_new)          # copies x_new into the static input, replays ~all launches with one driver call, returns the static output

``bench.py`` times the FLUX.1-schnell step this way (589-800 launches per replay).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor] = (), *, warmup: int = 2):
        """``fn(*tensors) -> tensor | tuple | None`` is called ``warmup`` times eagerly (lazy weight conversion, workspace allocation,
        per-kernel attribute setup must not happen inside the capture) and once under capture.  Inputs are cloned into static buffers."""
        if not torch.cuda.is_available():
            raise RuntimeError("nunchaku_b200 has no CPU path: GraphedStep needs a CUDA device")
        for t in example_inputs:
            if not (isinstance(t, torch.Tensor) and t.is_cuda):
                raise ValueError("example_inputs must be CUDA tensors")
        self.static_inputs = [t.clone() for t in example_inputs]
        self._fn = fn
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn(*self.static_inputs)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.static_output = fn(*self.static_inputs)
        cur.wait_stream(side)

    def replay(self) -> None:
        """Replays the captured launches on the current stream with whatever the static input buffers hold."""
        self.graph.replay()

    def __call__(self, *inputs: torch.Tensor):
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"expected {len(self.static_inputs)} inputs")
        for dst, src in zip(self.static_inputs, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError("inputs must keep the captured shapes and dtypes (a CUDA graph is shape-static)")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_output
