"""CPU emulation of the C-ABI kernels' CONTRACTS (include/erfnet_b200.h) in plain torch -- test infrastructure.

It lets the `-m "not gpu"` suite exercise the host logic of erfnet_pytorch_b200 (conv specs, weight packing for
fwd/dgrad, wgrad unpacking, BN bookkeeping, autograd wiring, DDP bucketing) against the oracle without a GPU.
The product never imports this; on a GPU box the same Python drives the real kernels (tests marked gpu)."""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

from erfnet_pytorch_b200 import ops


def _unpack(wp, spec):
    return wp.float().view(spec.kh, spec.kw, spec.cin_p, spec.w_ld)[:, :, :, :spec.cout]  # [kh,kw,ci,co]


def conv_fwd(x, wp, spec, bias=None, relu=False, out=None, out_coff=0, mask=None, add=None, stats=None, out_hw=None):
    n, h, w, ld = x.shape
    xi = x[..., :spec.cin_p].float().permute(0, 3, 1, 2)
    wt = _unpack(wp, spec)
    if out is not None:
        ho, wo = out.shape[1:3]
    else:
        ho, wo = out_hw if out_hw is not None else spec.out_hw(h, w)
    if not spec.transposed:
        y = F.conv2d(xi, wt.permute(3, 2, 0, 1), None, spec.stride, (spec.pad_h, spec.pad_w), (spec.dil_h, spec.dil_w))
    else:
        base_h = (h - 1) * spec.stride - 2 * spec.pad_h + spec.dil_h * (spec.kh - 1) + 1
        base_w = (w - 1) * spec.stride - 2 * spec.pad_w + spec.dil_w * (spec.kw - 1) + 1
        y = F.conv_transpose2d(xi, wt.permute(2, 3, 0, 1), None, spec.stride, (spec.pad_h, spec.pad_w), (ho - base_h, wo - base_w), 1,
                               (spec.dil_h, spec.dil_w))
    assert y.shape[2:] == (ho, wo), (y.shape, ho, wo)
    y = y.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias
    if stats is not None:
        c = stats.numel() // 2
        stats[out_coff:out_coff + spec.cout] += y.double().sum((0, 1, 2))
        stats[c + out_coff:c + out_coff + spec.cout] += (y.double() ** 2).sum((0, 1, 2))
    if add is not None:
        y = y + add[..., :spec.cout].float()
    if relu:
        y = y.relu()
    if mask is not None:
        y = y * (mask[..., :spec.cout].float() > 0)
    if out is None:
        out = torch.empty(n, ho, wo, spec.cout, dtype=x.dtype)
    out[..., out_coff:out_coff + spec.cout] = y.to(x.dtype)
    return out


@torch.enable_grad()
def conv_wgrad(x, dout, spec, dout_coff=0, dw_out=None, db_out=None, want_bias=True):
    n, h, w, _ = x.shape
    xi = x[..., :spec.cin_p].detach().float().permute(0, 3, 1, 2)
    wt = torch.zeros(spec.kh, spec.kw, spec.cin_p, spec.cout, requires_grad=True)
    dy = dout[..., dout_coff:dout_coff + spec.cout].detach().float().permute(0, 3, 1, 2)
    ho, wo = dout.shape[1:3]
    if not spec.transposed:
        y = F.conv2d(xi, wt.permute(3, 2, 0, 1), None, spec.stride, (spec.pad_h, spec.pad_w), (spec.dil_h, spec.dil_w))
    else:
        base_h = (h - 1) * spec.stride - 2 * spec.pad_h + spec.dil_h * (spec.kh - 1) + 1
        base_w = (w - 1) * spec.stride - 2 * spec.pad_w + spec.dil_w * (spec.kw - 1) + 1
        y = F.conv_transpose2d(xi, wt.permute(2, 3, 0, 1), None, spec.stride, (spec.pad_h, spec.pad_w), (ho - base_h, wo - base_w), 1,
                               (spec.dil_h, spec.dil_w))
    (g,) = torch.autograd.grad(y, wt, dy)   # [kh,kw,ci,co]
    g = g[:, :, :spec.cin]
    gt = g.permute(2, 3, 0, 1) if spec.transposed else g.permute(3, 2, 0, 1)
    dw = gt.contiguous() if dw_out is None else dw_out.add_(gt)
    db = None
    if want_bias:
        db = dy.sum((0, 2, 3)) if db_out is None else db_out.add_(dy.sum((0, 2, 3)))
    return dw, db


def maxpool_fwd(x, c, out, out_coff, scale=None, shift=None, relu=False, stats=None):
    m = F.max_pool2d(x[..., :c].float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    if stats is not None:
        ct = stats.numel() // 2
        stats[out_coff:out_coff + c] += m.double().sum((0, 1, 2))
        stats[ct + out_coff:ct + out_coff + c] += (m.double() ** 2).sum((0, 1, 2))
    if scale is not None:
        m = m * scale + shift
    if relu:
        m = m.relu()
    out[..., out_coff:out_coff + c] = m.to(out.dtype)


@torch.enable_grad()
def maxpool_bwd(x, c, dout, out_coff, dx):
    xi = x[..., :c].detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    m = F.max_pool2d(xi, 2, 2)
    (g,) = torch.autograd.grad(m, xi, dout[..., out_coff:out_coff + c].float().permute(0, 3, 1, 2))
    dx[..., :c] += g.permute(0, 2, 3, 1).to(dx.dtype)


def nchw_to_nhwc(x, ld, dtype):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, ld, dtype=dtype)
    out[..., :c] = x.permute(0, 2, 3, 1).to(dtype)
    return out


def nhwc_to_nchw(x, c):
    return x[..., :c].float().permute(0, 3, 1, 2).contiguous()


def affine_act(x, scale=None, shift=None, res=None, relu=True, out=None):
    y = x.float()
    if scale is not None:
        if scale.dim() == 2:
            y = y * scale[:, None, None, :] + shift[:, None, None, :]
        else:
            y = y * scale + shift
    if res is not None:
        y = y + res.float()
    if relu:
        y = y.relu()
    return y.to(x.dtype)


def channel_stats(x, stats):
    c = x.shape[-1]
    stats[:c] += x.double().sum((0, 1, 2))
    stats[c:] += (x.double() ** 2).sum((0, 1, 2))


def channel_sum(x, out):
    out += x.float().sum((0, 1, 2))


def channel_sum_slice(x, c, out):
    out += x[..., :c].float().sum((0, 1, 2))


def relu_bwd(g, mask):
    return (g.float() * (mask.float() > 0)).to(g.dtype)


def bn_finalize(stats, count, gamma, beta, rm, rv, momentum, eps, drop=None):
    c = gamma.numel()
    mu = stats[:c] / count
    var = (stats[c:] / count - mu * mu).clamp_min(0)
    stats.zero_()
    invstd = (1.0 / torch.sqrt(var + eps)).float()
    if rm is not None:
        unb = var * count / (count - 1) if count > 1 else var
        rm.mul_(1 - momentum).add_(momentum * mu.float())
        rv.mul_(1 - momentum).add_(momentum * unb.float())
    sc = gamma * invstd
    sh = beta - mu.float() * sc
    if drop is not None:
        sc, sh = sc[None, :] * drop, sh[None, :] * drop
    return mu.float(), invstd, sc.contiguous(), sh.contiguous()


def bn_eval_affine(gamma, beta, rm, rv, eps):
    sc = gamma / torch.sqrt(rv + eps)
    return sc, beta - rm * sc


def bn_bwd(g, y, gamma, mean, invstd, mask=None, drop=None, dgamma_out=None, dbeta_out=None, want_masked=False, beta_for_mask=None):
    # beta_for_mask: permission to recompute the mask from y (bf16 GPU kernel only); the mask tensor stays authoritative here
    if want_masked:
        gm = (g.float() * (mask.float() > 0) if mask is not None else g.float()).to(y.dtype)
        return bn_bwd(g, y, gamma, mean, invstd, mask, drop, dgamma_out, dbeta_out) + (gm,)
    gb = g.float()
    if mask is not None:
        gb = gb * (mask.float() > 0)
    if drop is not None:
        gb = gb * drop[:, None, None, :]
    xh = (y.float() - mean) * invstd
    m = y.shape[0] * y.shape[1] * y.shape[2]
    s1 = gb.double().sum((0, 1, 2)).float()
    s2 = (gb * xh).double().sum((0, 1, 2)).float()
    dy = gamma * invstd * (gb - s1 / m - xh * s2 / m)
    if dgamma_out is not None:
        dgamma_out.copy_(s2)
        dbeta_out.copy_(s1)
        return dy.to(y.dtype), dgamma_out, dbeta_out
    return dy.to(y.dtype), s2, s1


def ce_weight_sum(target, weight, c, ignore_index, wsum):
    valid = (target != ignore_index) & (target >= 0) & (target < c)    # the kernels skip EVERY label outside [0, C) (csrc/loss_metric.cu)
    t = torch.where(valid, target, torch.zeros_like(target))
    w = torch.ones(c) if weight is None else weight
    wsum += (w[t] * valid).sum()


@torch.enable_grad()
def ce_fwd_bwd(logits, target, weight, c, ignore_index, wsum, lsum, nchw=False, want_grad=True, colsum=None):
    assert colsum is None    # only the bf16 GPU kernel offers it (ops.ce_colsum_supported is False under emulation)
    lg = logits.detach().float() if nchw else logits.detach()[..., :c].float().permute(0, 3, 1, 2)
    lg = lg.requires_grad_(True)
    valid = (target != ignore_index) & (target >= 0) & (target < c)
    t = torch.where(valid, target, torch.zeros_like(target))
    w = torch.ones(c) if weight is None else weight
    wpx = w[t] * valid
    nll = -F.log_softmax(lg, 1).gather(1, t.unsqueeze(1)).squeeze(1)
    ls = (wpx * nll).sum()
    lsum += ls.detach()
    if not want_grad:
        return None
    (g,) = torch.autograd.grad(ls / wsum[0], lg)
    if nchw:
        return g
    dlog = torch.zeros_like(logits)
    dlog[..., :c] = g.permute(0, 2, 3, 1).to(logits.dtype)
    return dlog


def confusion_hist(pred, gt, nc, cm):
    p, g = pred.reshape(-1).long(), gt.reshape(-1).long()
    ok = (p >= 0) & (p < nc) & (g >= 0) & (g < nc)
    cm += torch.bincount(g[ok] * nc + p[ok], minlength=nc * nc).view(nc, nc)


def iou_counts(cm, nc, ignore, kept):
    tp = cm.diag()[:kept]
    rows = [g for g in range(nc) if g != ignore]
    fp = cm[rows].sum(0)[:kept] - tp
    fn = cm.sum(1)[:kept] - tp
    return torch.stack([tp, fp, fn]).double()


def argmax(logits, c, nchw, out_dtype=torch.int64):
    return (logits.float().argmax(1) if nchw else logits[..., :c].float().argmax(-1)).to(out_dtype)


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    gv = g * grad_scale + weight_decay * p
    m.mul_(beta1).add_(gv, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gv, gv, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale=1.0):
    step_dev += 1
    adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, int(step_dev), grad_scale)


def adam_step_graph(p, g, m, v, hyper_dev, beta1, beta2, eps, step_dev, grad_scale=1.0):
    step_dev += 1
    adam_step(p, g, m, v, float(hyper_dev[0]), beta1, beta2, eps, float(hyper_dev[1]), int(step_dev), grad_scale)


def sgd_step_graph(p, g, buf, hyper_dev, momentum, grad_scale=1.0):
    sgd_step(p, g, buf, float(hyper_dev[0]), momentum, float(hyper_dev[1]), False, grad_scale)


def _ml_per_px(kind, a, b, prob, target, eps):
    """per-pixel loss of erf_masked_loss_* (include/erfnet_b200.h) on flat fp32 operands, shape of target"""
    t = target
    if kind == ops.LOSS_L1:
        return (a.reshape(t.shape) - t).abs()
    if kind == ops.LOSS_LOGLIK:
        v = b.reshape(t.shape) + eps
        return (2 * math.pi * v).log() + (a.reshape(t.shape) - t) ** 2 / v
    if kind == ops.LOSS_MSE_WEIGHTED:
        return (a.reshape(t.shape) - t) ** 2
    c = prob.shape[1]
    tt = t.reshape(prob.shape[0], 1, prob.shape[2], prob.shape[3])
    d = a.reshape(1, c, 1, 1) - tt
    if kind == ops.LOSS_L1_CLASSPROB:
        val = d.abs()
    else:
        v = b.reshape(1, c, 1, 1) + eps
        val = (2 * math.pi * v).log() + d * d / v
    return (val * prob).sum(1).reshape(t.shape)


def masked_loss_fwd(kind, a, b, prob, target, eps, acc):
    valid = torch.ones_like(target, dtype=torch.bool) if kind == ops.LOSS_MSE_WEIGHTED else target != -1.0
    per = _ml_per_px(kind, a, b, prob, target, eps)
    acc[0] += (per * valid).double().sum()
    acc[1] += valid.double().sum()


@torch.enable_grad()
def masked_loss_bwd(kind, a, b, prob, target, eps, scale, acc, go):
    target = target.detach()
    valid = torch.ones_like(target, dtype=torch.bool) if kind == ops.LOSS_MSE_WEIGHTED else target != -1.0
    leaves = [None if t is None else t.detach().clone().requires_grad_(True) for t in (a, b, prob)]
    per = _ml_per_px(kind, leaves[0], leaves[1], leaves[2], target, eps)
    g = float(go.reshape(-1)[0]) * scale / float(acc[1])
    if prob is None:
        ((per * valid).sum() * g).backward()
        return leaves[0].grad.detach(), None if b is None else leaves[1].grad.detach()
    (per * valid).sum().backward()  # class sums stay unscaled
    c = prob.shape[1]
    dclass = torch.zeros(2 * c, dtype=torch.float64)
    dclass[:c] = leaves[0].grad.reshape(-1).double()
    if b is not None:
        dclass[c:] = leaves[1].grad.reshape(-1).double()
    return (leaves[2].grad * g).detach(), dclass


def softmax_nhwc_fwd(logits, c, drop=None):
    v = logits[..., :c].float()
    if drop is not None:
        v = v * drop[:, None, None, :]
    return torch.softmax(v, dim=-1).permute(0, 3, 1, 2).contiguous()


def softmax_nhwc_bwd(prob, dprob, drop, ld, dtype):
    p, dp = prob.permute(0, 2, 3, 1), dprob.float().permute(0, 2, 3, 1)
    g = p * (dp - (p * dp).sum(-1, keepdim=True))
    if drop is not None:
        g = g * drop[:, None, None, :]
    out = torch.zeros(*g.shape[:3], ld, dtype=dtype)
    out[..., :g.shape[-1]] = g.to(dtype)
    return out


def avgpool_fwd(x, c, k):
    return F.avg_pool2d(x[..., :c].float().permute(0, 3, 1, 2), k, 1, 0).permute(0, 2, 3, 1).contiguous()


@torch.enable_grad()
def avgpool_bwd(dout, h, w, ld, k, dtype):
    n, _, _, c = dout.shape
    xi = torch.zeros(n, c, h, w, requires_grad=True)
    (g,) = torch.autograd.grad(F.avg_pool2d(xi, k, 1, 0), xi, dout.permute(0, 3, 1, 2))
    dx = torch.zeros(n, h, w, ld, dtype=dtype)
    dx[..., :c] = g.permute(0, 2, 3, 1).to(dtype)
    return dx


def linear_fwd(x, w, bias):
    return F.linear(x, w, bias)


def linear_bwd(x, w, dout, need_dx=True, dw_out=None, db_out=None, want_bias=True):
    dx = dout @ w if need_dx else None
    dw = dout.t() @ x
    dw = dw if dw_out is None else dw_out.add_(dw)
    db = None
    if want_bias:
        db = dout.sum(0) if db_out is None else db_out.add_(dout.sum(0))
    return dx, dw, db


def ce_rows_fwd(logits, target, ignore_index):
    n, c = logits.shape
    valid = (target != ignore_index) & (target >= 0) & (target < c)
    lse = torch.logsumexp(logits, 1)
    t = target.clamp(0, c - 1)
    lt = logits.gather(1, t[:, None])[:, 0]
    nll = (lse - lt)[valid]
    idx = torch.arange(c)[None, :]
    rank = ((logits > lt[:, None]) | ((logits == lt[:, None]) & (idx < t[:, None]))).sum(1).int()
    rank[~valid] = -1
    return torch.stack([nll.sum(), valid.sum().float()]), lse, rank


def ce_rows_bwd(logits, target, ignore_index, lse, acc, go):
    n, c = logits.shape
    valid = (target != ignore_index) & (target >= 0) & (target < c)
    g = torch.exp(logits - lse[:, None])
    g[valid, target[valid]] -= 1.0
    g[~valid] = 0.0
    return g * ((go if go is not None else 1.0) / acc[1].clamp_min(1.0))


def sgd_step(p, g, buf, lr, momentum, weight_decay, first_step, grad_scale=1.0):
    d = g * grad_scale + weight_decay * p
    if momentum != 0:
        if first_step:
            buf.copy_(d)
        else:
            buf.mul_(momentum).add_(d)
        d = buf
    p.sub_(lr * d)


def ema_update(teacher, student, rate):
    teacher.mul_(1.0 - rate).add_(rate * student)


def ingest_u8(img, ld, dtype, mean=None, std=None):
    n, h, w, _ = img.shape
    v = img.float() / 255.0
    if mean is not None or std is not None:
        v = (v - (mean if mean is not None else 0.0)) / (std if std is not None else 1.0)
    out = torch.zeros(n, h, w, ld, dtype=dtype)
    out[..., :3] = v.to(dtype)
    return out


def label_lut(labels, lut):
    v = labels.long()
    inside = (v >= 0) & (v < lut.numel())
    return torch.where(inside, lut[v.clamp(0, lut.numel() - 1)], v)


def colorize(labels, cmap):
    v = labels.long()
    inside = (v >= 0) & (v < cmap.shape[0])
    col = cmap[v.clamp(0, cmap.shape[0] - 1)] * inside[..., None].to(torch.uint8)
    return col.permute(0, 3, 1, 2).contiguous()


_NAMES = ["conv_fwd", "conv_wgrad", "maxpool_fwd", "maxpool_bwd", "nchw_to_nhwc", "nhwc_to_nchw", "affine_act", "channel_stats", "channel_sum", "channel_sum_slice", "relu_bwd", "bn_finalize",
          "bn_eval_affine", "bn_bwd", "ce_weight_sum", "ce_fwd_bwd", "confusion_hist", "iou_counts", "argmax", "adam_step", "adam_step_dev", "adam_step_graph", "sgd_step_graph", "masked_loss_fwd", "masked_loss_bwd", "softmax_nhwc_fwd", "softmax_nhwc_bwd",
          "avgpool_fwd", "avgpool_bwd", "linear_fwd", "linear_bwd", "ce_rows_fwd", "ce_rows_bwd", "sgd_step", "ema_update", "ingest_u8", "label_lut", "colorize"]


# ---- optional: contracts of the tcgen05 3-tap conv entry points, so that the HOST wiring of the tensor-core path (operand images for forward and
# data gradient, epilogue options, statistics / column sums, weight-gradient accumulation, the folded 32-channel block) runs on the CPU too.
# An "image" here is simply Wt[t][co][ci] in fp32 (the real one is the same matrix in bf16, swizzled for shared memory).
def _tc_image(w, dgrad=False):
    co, ci = w.shape[:2]
    wt = w.detach().float().reshape(co, ci, 3).permute(2, 0, 1)      # [t][co][ci]
    if dgrad:
        wt = wt.flip(0).transpose(1, 2)                              # taps mirrored, channels transposed
    return wt.contiguous()


def tc_conv1d_tc_supported(c, w, dil, dtype):
    return c in (16, 64, 128) and w >= 16 and 1 <= dil <= 32


def tc_pack_conv1d_tc(w, dgrad=False):
    return _tc_image(w, dgrad)


def tc_pack_conv1d_tc_pair(w):
    return _tc_image(w, False), _tc_image(w, True)


def tc_conv1d_tc(x, w_img, bias, horizontal, dil, relu=False, res=None, mask=None, stats=None, colsum=None, static_w=False):
    assert not (res is not None and mask is not None)
    c = x.shape[-1]
    xi = x.float().permute(0, 3, 1, 2)
    wt = w_img.permute(1, 2, 0)                                       # [co][ci][t]
    wt = wt.unsqueeze(2) if horizontal else wt.unsqueeze(3)           # [co,ci,1,3] / [co,ci,3,1]
    y = F.conv2d(xi, wt, None, 1, (0, dil) if horizontal else (dil, 0), (1, dil) if horizontal else (dil, 1)).permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias
    if res is not None:
        y = y + res.float()
    if relu:
        y = y.relu()
    if mask is not None:
        y = y * (mask.float() > 0)
    out = y.to(x.dtype)
    if stats is not None:      # statistics of the STORED output
        stats[:c] += out.double().sum((0, 1, 2))
        stats[c:] += (out.double() ** 2).sum((0, 1, 2))
    if colsum is not None:
        colsum += out.float().sum((0, 1, 2))
    return out


def tc_conv1d_wgrad_tc_supported(c, w, dil, dtype):
    return c in (16, 64, 128) and w >= 16 and 1 <= dil <= 32


@torch.enable_grad()
def tc_conv1d_wgrad_tc(x, dy, horizontal, dil, dw_out=None, db_out=None, want_bias=True):
    c = x.shape[-1]
    xi = x.detach().float().permute(0, 3, 1, 2)
    wt = torch.zeros((c, c, 1, 3) if horizontal else (c, c, 3, 1), requires_grad=True)
    y = F.conv2d(xi, wt, None, 1, (0, dil) if horizontal else (dil, 0), (1, dil) if horizontal else (dil, 1))
    (g,) = torch.autograd.grad(y, wt, dy.detach().float().permute(0, 3, 1, 2))
    dw = g if dw_out is None else dw_out.add_(g)
    db = None
    if want_bias:
        s = dy.float().sum((0, 1, 2))
        db = s if db_out is None else db_out.add_(s)
    return dw, db


_TC_NAMES = ["conv1d_tc_supported", "pack_conv1d_tc", "pack_conv1d_tc_pair", "conv1d_tc", "conv1d_wgrad_tc_supported", "conv1d_wgrad_tc"]


@contextlib.contextmanager
def emulated_kernels(tc=False):
    """Swap erfnet_pytorch_b200.ops kernels for the CPU emulations above (tests only).  tc=True also emulates the tcgen05 3-tap conv entry
    points, so blocks take their tensor-core code path (at widths that are multiples of 128) instead of the generic one."""
    names = _NAMES + (_TC_NAMES if tc else [])
    saved = {n: getattr(ops, n) for n in names}
    saved["EMULATED"] = ops.EMULATED
    try:
        ops.EMULATED = True
        for n in _NAMES:
            setattr(ops, n, globals()[n])
        if tc:
            for n in _TC_NAMES:
                setattr(ops, n, globals()["tc_" + n])
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
