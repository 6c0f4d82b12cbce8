"""CPU oracle for the ERFNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``erfnet_pytorch_b200``) never does and has no CPU fallback.

What it is: a functional, state_dict-driven restatement of the reference's arithmetic in
plain fp32 torch ops on the CPU (the reference's arithmetic *is* torch ops; nothing here is
copied -- each function re-derives the dataflow from the cited reference lines and takes the
weights from a flat ``{key: tensor}`` dict with the reference's checkpoint key names).

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference modules themselves, executed in the build
container by ``oracle/make_golden.py`` (which imports /root/reference by path) and committed
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the oracle against
them on every CPU run.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]
BN_EPS = 1e-3  # eval/erfnet.py:18,34,40,103


# ----------------------------------------------------------------------------------------------
# deterministic weights (independent of module construction order, so that the reference, the
# oracle and the CUDA modules can all be given the same values without sharing RNG state)
# ----------------------------------------------------------------------------------------------
def erfnet_param_shapes(num_classes: int = 20) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape for every entry of ``eval/erfnet.py::ERFNet(num_classes).state_dict()``
    (345 entries for 20 classes)."""
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)
        shapes[p + ".running_mean"] = (c,)
        shapes[p + ".running_var"] = (c,)
        shapes[p + ".num_batches_tracked"] = ()

    def down(p, ci, co):  # eval/erfnet.py:12-18
        shapes[p + ".conv.weight"] = (co - ci, ci, 3, 3)
        shapes[p + ".conv.bias"] = (co - ci,)
        bn(p + ".bn", co)

    def nb(p, c):  # eval/erfnet.py:26-42
        for name, k in (("conv3x1_1", (3, 1)), ("conv1x3_1", (1, 3))):
            shapes[f"{p}.{name}.weight"] = (c, c) + k
            shapes[f"{p}.{name}.bias"] = (c,)
        bn(p + ".bn1", c)
        for name, k in (("conv3x1_2", (3, 1)), ("conv1x3_2", (1, 3))):
            shapes[f"{p}.{name}.weight"] = (c, c) + k
            shapes[f"{p}.{name}.bias"] = (c,)
        bn(p + ".bn2", c)

    def up(p, ci, co):  # eval/erfnet.py:99-103
        shapes[p + ".conv.weight"] = (ci, co, 3, 3)
        shapes[p + ".conv.bias"] = (co,)
        bn(p + ".bn", co)

    down("encoder.initial_block", 3, 16)
    down("encoder.layers.0", 16, 64)
    for i in range(1, 6):
        nb(f"encoder.layers.{i}", 64)
    down("encoder.layers.6", 64, 128)
    for i in range(7, 15):
        nb(f"encoder.layers.{i}", 128)
    shapes["encoder.output_conv.weight"] = (num_classes, 128, 1, 1)
    shapes["encoder.output_conv.bias"] = (num_classes,)
    up("decoder.layers.0", 128, 64)
    nb("decoder.layers.1", 64)
    nb("decoder.layers.2", 64)
    up("decoder.layers.3", 64, 16)
    nb("decoder.layers.4", 16)
    nb("decoder.layers.5", 16)
    shapes["decoder.output_conv.weight"] = (16, num_classes, 2, 2)
    shapes["decoder.output_conv.bias"] = (num_classes,)
    return shapes


def _key_seed(seed: int, key: str) -> int:
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFF


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, gain: float = 1.0) -> StateDict:
    """Deterministic "trained-like" weights: conv weights ~ U(+-gain*sqrt(3/fan_in)) (keeps activations
    O(1) through 70 layers), small biases, BN gamma in [0.7,1.3], beta in +-0.2, running_mean +-0.3,
    running_var in [0.5,1.5].  Values depend only on (seed, key)."""
    sd: StateDict = {}
    for key, shape in shapes.items():
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
                fan_in = shape[0] * shape[2] * shape[3] / 4.0  # ConvTranspose, stride 2: ~1/4 of the taps hit
            elif len(shape) == 2:       # nn.Linear [out, in]
                fan_in = shape[1]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            bound = gain * math.sqrt(3.0 / fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


# ----------------------------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------------------------
def _bn(x, sd, p, training, momentum=0.1):
    """nn.BatchNorm2d(C, eps=1e-3): train = batch mean / biased var, running stats updated in place with
    momentum 0.1 and the unbiased var; eval = running stats (eval/erfnet.py:18,22)."""
    if training:
        sd[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training, momentum, BN_EPS)


def downsampler_block(x, sd: StateDict, p: str, training=False, use_batchnorm=True):
    """eval/erfnet.py:20-23 (gated BN: train/erfnet_blocks.py:24-29): cat([conv3x3 s2 p1, maxpool2x2]) -> BN -> ReLU;
    conv channels first, pooled input channels after."""
    y = torch.cat([F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=1),
                   F.max_pool2d(x, 2, stride=2)], 1)
    if use_batchnorm:
        y = _bn(y, sd, p + ".bn", training)
    return F.relu(y)


def non_bottleneck_1d(x, sd: StateDict, p: str, dilated: int, dropprob: float = 0.0, training=False,
                      drop_scale: Optional[torch.Tensor] = None, use_batchnorm=True):
    """eval/erfnet.py:45-61.  ``drop_scale`` is the Dropout2d multiplier [N,C,1,1] (0 or 1/(1-p)); when None and
    training with dropprob != 0 it is drawn with torch RNG exactly like nn.Dropout2d (noise shape [N,C,1,1])."""
    y = F.relu(F.conv2d(x, sd[p + ".conv3x1_1.weight"], sd[p + ".conv3x1_1.bias"], padding=(1, 0)))
    y = F.conv2d(y, sd[p + ".conv1x3_1.weight"], sd[p + ".conv1x3_1.bias"], padding=(0, 1))
    if use_batchnorm:
        y = _bn(y, sd, p + ".bn1", training)
    y = F.relu(y)
    y = F.relu(F.conv2d(y, sd[p + ".conv3x1_2.weight"], sd[p + ".conv3x1_2.bias"], padding=(dilated, 0),
                        dilation=(dilated, 1)))
    y = F.conv2d(y, sd[p + ".conv1x3_2.weight"], sd[p + ".conv1x3_2.bias"], padding=(0, dilated), dilation=(1, dilated))
    if use_batchnorm:
        y = _bn(y, sd, p + ".bn2", training)
    if dropprob != 0 and training:  # eval/erfnet.py:58-59
        if drop_scale is None:
            y = F.dropout2d(y, dropprob, True)
        else:
            y = y * drop_scale
    return F.relu(y + x)


def upsampler_block(x, sd: StateDict, p: str, training=False, use_batchnorm=True):
    """eval/erfnet.py:105-108: ConvTranspose2d(3, s2, p1, op1) -> BN -> ReLU."""
    y = F.conv_transpose2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=1, output_padding=1)
    if use_batchnorm:
        y = _bn(y, sd, p + ".bn", training)
    return F.relu(y)


ENC_DILATIONS = (2, 4, 8, 16, 2, 4, 8, 16)  # eval/erfnet.py:78-82


def encoder_forward(x, sd: StateDict, predict=False, training=False, drop_scales=None, p="encoder", dropprob=0.1):
    """eval/erfnet.py:87-96.  ``drop_scales``: optional {layer_prefix: [N,C,1,1]} of Dropout2d multipliers."""
    ds = drop_scales or {}
    y = downsampler_block(x, sd, p + ".initial_block", training)
    y = downsampler_block(y, sd, p + ".layers.0", training)
    for i in range(1, 6):
        q = f"{p}.layers.{i}"
        y = non_bottleneck_1d(y, sd, q, 1, dropprob, training, ds.get(q))
    y = downsampler_block(y, sd, p + ".layers.6", training)
    for i, d in enumerate(ENC_DILATIONS):
        q = f"{p}.layers.{7 + i}"
        y = non_bottleneck_1d(y, sd, q, d, dropprob, training, ds.get(q))
    if predict:
        y = F.conv2d(y, sd[p + ".output_conv.weight"], sd[p + ".output_conv.bias"])
    return y


def decoder_forward(x, sd: StateDict, training=False, p="decoder"):
    """eval/erfnet.py:126-134."""
    y = upsampler_block(x, sd, p + ".layers.0", training)
    y = non_bottleneck_1d(y, sd, p + ".layers.1", 1, 0.0, training)
    y = non_bottleneck_1d(y, sd, p + ".layers.2", 1, 0.0, training)
    y = upsampler_block(y, sd, p + ".layers.3", training)
    y = non_bottleneck_1d(y, sd, p + ".layers.4", 1, 0.0, training)
    y = non_bottleneck_1d(y, sd, p + ".layers.5", 1, 0.0, training)
    return F.conv_transpose2d(y, sd[p + ".output_conv.weight"], sd[p + ".output_conv.bias"], stride=2)


def erfnet_forward(x, sd: StateDict, only_encode=False, training=False, drop_scales=None, dropprob=0.1):
    """eval/erfnet.py:147-152.  ``dropprob`` = encoder Dropout2d p (0.1 in the reference, eval/erfnet.py:74,79-82)."""
    if only_encode:
        return encoder_forward(x, sd, predict=True, training=training, drop_scales=drop_scales, dropprob=dropprob)
    return decoder_forward(encoder_forward(x, sd, training=training, drop_scales=drop_scales, dropprob=dropprob), sd, training)


# ----------------------------------------------------------------------------------------------
# fork: ladder regression net (train/net_regression_ladder.py; BN off, dropout off for softmax_classes = 0)
# ----------------------------------------------------------------------------------------------
def decoder_block(x, sd: StateDict, p: str, n_nb: int = 2):
    """train/erfnet_blocks.py:114-135: UpsamplerBlock -> n_nb x non_bottleneck_1d(C, 0, 1), BatchNorm off."""
    y = upsampler_block(x, sd, p + ".layers.0", use_batchnorm=False)
    for i in range(n_nb):
        y = non_bottleneck_1d(y, sd, f"{p}.layers.{i + 1}", 1, use_batchnorm=False)
    return y


def ladder_block(x, ladder, sd: StateDict, p: str):
    """train/erfnet_blocks.py:140-154: cat -> non_bottleneck_1d(2C, 0.03, 1) (dropout off) -> conv3x3 p1 -> ReLU."""
    y = torch.cat((x, ladder), dim=1)
    y = non_bottleneck_1d(y, sd, p + ".layer", 1, use_batchnorm=False)
    return F.relu(F.conv2d(y, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1))


def regression_ladder_forward(x, sd: StateDict):
    """train/net_regression_ladder.py:41-52, 72-80, 122-128 with softmax_classes = 0: [B, 1 or 2, H, W]."""
    e1 = downsampler_block(x, sd, "encoder.initial_block", use_batchnorm=False)
    e2 = downsampler_block(e1, sd, "encoder.block1.0", use_batchnorm=False)
    for i in range(1, 6):
        e2 = non_bottleneck_1d(e2, sd, f"encoder.block1.{i}", 1, use_batchnorm=False)
    e3 = downsampler_block(e2, sd, "encoder.block2.0", use_batchnorm=False)
    for i, d in enumerate(ENC_DILATIONS):
        e3 = non_bottleneck_1d(e3, sd, f"encoder.block2.{i + 1}", d, use_batchnorm=False)
    y = decoder_block(e3, sd, "decoder.scalar_decoder_1")
    y = ladder_block(y, e2, sd, "decoder.ladder_block_1")
    y = decoder_block(y, sd, "decoder.scalar_decoder_2")
    y = ladder_block(y, e1, sd, "decoder.ladder_block_2")
    return F.conv_transpose2d(y, sd["decoder.scalar_output_conv.weight"], sd["decoder.scalar_output_conv.bias"], stride=2)


# ----------------------------------------------------------------------------------------------
# loss, metric, optimiser
# ----------------------------------------------------------------------------------------------
def cross_entropy_2d(logits, target, weight=None, ignore_index=-1):
    """train/evaluation_functions.py:6-14: nn.CrossEntropyLoss(ignore_index=-1) on [B,C,H,W] logits and int64
    [B,H,W] targets = sum_valid(w[y]*nll)/sum_valid(w[y]).  The reference drops ``weight``; the extension honours
    it (weight=None is the fork-exact behaviour)."""
    logp = F.log_softmax(logits.float(), dim=1)
    valid = target != ignore_index
    t = torch.where(valid, target, torch.zeros_like(target))
    nll = -logp.gather(1, t.unsqueeze(1)).squeeze(1)
    w = torch.ones_like(nll) if weight is None else weight.float()[t]
    w = w * valid
    return (w * nll).sum() / w.sum()


def _masked_mean(per_px, target):
    """mean of per_px over the labelled pixels (target != -1): masked_select(...).mean() of the reference."""
    m = target != -1.0
    return (per_px * m).sum() / m.sum()


def l1_loss_masked(outputs, targets):
    """train/evaluation_functions.py:89-97: outputs [B,1,H,W], targets float [B,H,W]."""
    return _masked_mean((outputs.squeeze(1) - targets).abs(), targets)


def loglik_loss_masked(cost, var, targets, opt_eps=0.0):
    """train/evaluation_functions.py:71-87 (the 1/2 factor of the Gaussian likelihood is dropped there too)."""
    v = var + opt_eps
    return _masked_mean((2 * math.pi * v).log() + (cost - targets) ** 2 / v, targets)


def l1_loss_classprob_masked(prob, cost, targets):
    """train/evaluation_functions.py:32-47: prob [B,C,H,W], cost [1,C,1,1], targets [B,H,W]."""
    t = targets.unsqueeze(1)
    return _masked_mean(((cost - t).abs() * prob).sum(1), targets)


def loglik_loss_classprob_masked(prob, cost, var, targets, opt_eps=0.0):
    """train/evaluation_functions.py:49-69."""
    t = targets.unsqueeze(1)
    v = var + opt_eps
    return _masked_mean((((2 * math.pi * v).log() + (cost - t) ** 2 / v) * prob).sum(1), targets)


def mse_loss_weighted(outputs, targets, scale_factor, weight):
    """train/evaluation_functions.py:122-131."""
    return ((outputs - targets) ** 2 * scale_factor * weight).mean()


def accuracy_counts(prob, targets, n_classes):
    """train/evaluation_functions.py:133-163 on integers: (n_correct, n_total, per-class correct, per-class total)."""
    pred = prob.argmax(1)
    correct = pred == targets
    cc = np.array([int((correct & (targets == c)).sum()) for c in range(n_classes)], np.int64)
    ct = np.array([int((targets == c).sum()) for c in range(n_classes)], np.int64)
    return int(correct.sum()), int((targets >= 0).sum()), cc, ct


def iou_counts(pred: np.ndarray, gt: np.ndarray, n_classes: int = 20, ignore_index: int = 19):
    """eval/iouEval.py:21-69 restated on integers.  pred, gt: integer label maps of identical shape.
    tp[c] = #(p==c & g==c); fp[c] = #(p==c & g!=c & g!=ignore); fn[c] = #(p!=c & g==c) for the kept classes
    (ignore channel dropped when n_classes > ignore_index, eval/iouEval.py:12,48-51)."""
    ign = ignore_index if n_classes > ignore_index else -1
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    kept = n_classes if ign == -1 else n_classes - 1
    tp = np.zeros(kept, np.int64)
    fp = np.zeros(kept, np.int64)
    fn = np.zeros(kept, np.int64)
    for c in range(kept):
        # the reference slices channels [:ignoreIndex], which equals "all but the last" only for ignore==n-1;
        # keep its exact slice semantics
        if ign != -1 and c >= ign:
            break
        pc = pred == c
        gc = gt == c
        tp[c] = np.count_nonzero(pc & gc)
        fp[c] = np.count_nonzero(pc & ~gc & (gt != ign)) if ign != -1 else np.count_nonzero(pc & ~gc)
        fn[c] = np.count_nonzero(~pc & gc)
    return tp, fp, fn


def iou_from_counts(tp, fp, fn):
    """eval/iouEval.py:71-75 (float64 accumulators)."""
    tp = np.asarray(tp, np.float64)
    den = tp + np.asarray(fp, np.float64) + np.asarray(fn, np.float64) + 1e-15
    iou = tp / den
    return float(iou.mean()), iou


def adam_step(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-4, weight_decay=1e-6):
    """torch.optim.Adam as configured by train/main.py:43-46,213 (L2 folded into the gradient, not AdamW).
    In-place on p, m, v (fp32 tensors); ``step`` is the 1-based step count."""
    g = g + weight_decay * p
    m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
    return p
