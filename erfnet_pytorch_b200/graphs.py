"""CUDA-graph wrappers: the hot loops are sequences of ~90 (inference) / ~700 (training) short kernels, so at batch
sizes of a few images the CPU launch path, not the GPU, sets the pace.  Capturing the sequence once and replaying it
removes that.  Training graphs live in parallel.FlatAdamTrainer.capture_cuda_graph."""
from __future__ import annotations

import torch


class GraphedInference:
    """Eval-mode forward of a model for one fixed input shape and the weights it had at capture time (the graph holds the packed,
    BN-folded weight images: build a new GraphedInference after load_state_dict() / an optimiser step).  ``__call__(x)`` copies x into a
    static buffer, replays the captured graph and returns the static output tensor (valid until the next call)."""

    def __init__(self, model, example: torch.Tensor, warmup: int = 2, **fwd_kwargs):
        model.eval()
        self.model = model
        self.x = example.clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model(self.x, **fwd_kwargs)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = model(self.x, **fwd_kwargs)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
