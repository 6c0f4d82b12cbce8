"""Deterministic synthetic "trained-like" weights for benchmarks and profiling targets (the shipped pretrained blob is missing from the
reference checkout).  Values depend only on (seed, checkpoint key): conv weights ~ U(+-gain*sqrt(3/fan_in)) (keeps
activations O(1) through 70 layers), small biases, BN gamma in [0.7, 1.3], beta in +-0.2, running_mean +-0.3, running_var in [0.5, 1.5].
This is input generation, not arithmetic of the hot path; the test oracle carries its own copy (oracle/erfnet_oracle.make_state_dict) and
tests/test_host_logic.py checks that the two agree bit for bit."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch


def _key_seed(seed: int, key: str) -> int:
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFF


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    sd = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        g = torch.Generator().manual_seed(_key_seed(seed, key))
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(8, dtype=torch.int64)
        elif key.endswith("running_mean"):
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * 0.6
        elif key.endswith("running_var"):
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif ".bn" in key and key.endswith(".weight"):
            sd[key] = 0.7 + 0.6 * torch.rand(shape, generator=g)
        elif ".bn" in key and key.endswith(".bias"):
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * 0.4
        elif key.endswith(".bias") or len(shape) == 1:
            sd[key] = (torch.rand(shape, generator=g) - 0.5) * 0.2
        else:
            if "decoder" in key and ".conv." in key or key.startswith("decoder.output_conv"):
                fan_in = shape[0] * shape[2] * shape[3] / 4.0      # ConvTranspose, stride 2: ~1/4 of the taps hit
            elif len(shape) == 2:                                   # nn.Linear [out, in]
                fan_in = shape[1]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * (gain * math.sqrt(3.0 / fan_in))
    return sd


def synthetic_state_dict(model: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """make_state_dict for the checkpoint layout of ``model``."""
    return make_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
