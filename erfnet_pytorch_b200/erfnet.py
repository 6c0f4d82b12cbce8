"""B200-native ERFNet with the reference's module API (eval/erfnet.py): same class names, constructor
signatures, attribute names and fp32 state_dict layout (345 entries for 20 classes), NCHW fp32 tensors at the
boundary -- but every arithmetic op runs in liberfnet_b200.so on NHWC bf16 (or fp32 "check mode") activations.

Differences from the reference that a caller can observe:
  * CUDA only; there is no CPU / eager fallback (RuntimeError instead).
  * ``model.precision`` ("bf16" default, "fp32" check mode) selects activation storage; accumulation is fp32.
  * in ``eval()`` mode no autograd graph is built (the reference builds one it never uses, eval/eval_iou.py:89);
    in ``train()`` mode outputs carry a grad_fn whose backward runs this library's dgrad/wgrad kernels.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ConvSpec

BN_EPS = 1e-3
BN_MOMENTUM = 0.1
_DEFAULT_PRECISION = "bf16"


def set_default_precision(p: str):
    global _DEFAULT_PRECISION
    assert p in ("bf16", "fp32")
    _DEFAULT_PRECISION = p


def _act_dtype(precision):
    return torch.bfloat16 if precision == "bf16" else torch.float32


class _Ctx:
    """Per-forward context handed down the module tree (precision, dropout masks, weight cache)."""

    def __init__(self, precision, training, drop_scales=None):
        self.dtype = _act_dtype(precision)
        self.training = training
        self.drop_scales = drop_scales  # optional {id(module) or name: [N,C] fp32} (tests inject fixed masks)
        self.drawn = None               # {module: [N,C] fp32} Dropout2d multipliers drawn up front by draw_dropout()


# ---------------------------------------------------------------------------------------------------------------
# autograd glue: one Function per block; the block object supplies fwd_train()/bwd()
# ---------------------------------------------------------------------------------------------------------------
class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, block, bctx, x, *params):
        y, saved = block._fwd_train(x.detach(), bctx)
        # detached aliases: keeping `y` itself would create a node -> y -> grad_fn(node) reference cycle
        ctx.block, ctx.saved = block, tuple(t.detach() if isinstance(t, torch.Tensor) and t.requires_grad else t for t in saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, grads = ctx.block._bwd(dy.contiguous(), ctx.saved, ctx.needs_input_grad[2])
        if _GRAD_SINK is not None:
            ctx.saved = None      # the flat trainer runs exactly one backward per forward: release the activations block by block.  Plain
                                  # autograd use keeps them until the graph dies, so backward(retain_graph=True) + a second backward
                                  # (train/main.py:375-379, mean teacher) works
        return (None, None, dx) + tuple(grads)


class _Holder:
    """(weight, bias) of a derived conv: quacks like the nn.Conv2d the block code reads them from."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _FoldAdapter:
    """Presents a folded 32-channel block to _BlockFn: the forward / backward kernels run on its 64-channel twin with folded weights, the
    parameter gradients are mapped back (and accumulated into the gradient sink when one is active), all without native autograd nodes."""

    def __init__(self, block, twin):
        self.block, self.twin = block, twin

    def _train_params(self):
        return self.block._train_params()

    def _fwd_train(self, x, ctx):
        self.twin.__dict__["_conv_override"] = [_Holder(a, b) for a, b in self.block._folded_weights()]
        return self.twin._fwd_train(x, ctx)

    def _bwd(self, dy, saved, need_dx):
        dx, g64 = self.twin._bwd(dy, saved, need_dx)                  # twin order: w1,b1,w2,b2,g1,be1,w3,b3,w4,b4,g2,be2 (all returned: no sink on holders)
        P = self.block._train_params()
        grads = [None] * 12
        for i, (wi, bi) in enumerate(((0, 1), (2, 3), (6, 7), (8, 9))):
            dw, db = self.block._unfold_grads(i, g64[wi], g64[bi])
            for idx, g in ((wi, dw), (bi, db)):
                dst = _sink(P[idx])
                if dst is not None:
                    dst.add_(g)
                else:
                    grads[idx] = g
        _sunk(*P)
        return dx, tuple(grads)


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ld, dtype):
        ctx.c = x.shape[1]
        return ops.nchw_to_nhwc(x, ld, dtype)

    @staticmethod
    def backward(ctx, dy):
        return ops.nhwc_to_nchw(dy.contiguous(), ctx.c), None, None


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c):
        ctx.ld, ctx.dtype = x.shape[-1], x.dtype
        return ops.nhwc_to_nchw(x, c)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_nhwc(dy, ctx.ld, ctx.dtype), None


def input_to_nhwc(input, ld, dtype, mean=None, std=None):
    """Public-boundary input -> internal NHWC activations.  NCHW fp32 (the reference API) goes through the layout kernel; a uint8
    [N,H,W,3] image batch (HWC, what PIL / a camera delivers; train/dataset.py:113-114 applies ToTensor to it) goes through the ingest
    kernel, which fuses ToTensor (/255) [+ Normalize] with the layout change."""
    if input.dtype == torch.uint8:
        return ops.ingest_u8(input, ld, dtype, mean, std)
    return _ToNHWC.apply(input, ld, dtype)


class _B200Module(nn.Module):
    """Common plumbing: NCHW<->NHWC at the public boundary, packed-weight cache for eval mode."""
    precision: Optional[str] = None
    _in_ld = None   # NHWC pitch of the input this block expects (None: = channels)
    _out_c = None   # logical channels of the output

    def _precision(self):
        return self.precision or _DEFAULT_PRECISION

    def _cache(self, dtype):
        """Folded + packed eval-mode operands, rebuilt when any parameter/buffer changes (load_state_dict, optimizer)."""
        sig = (dtype,) + _sig(self)
        c = self.__dict__.get("_packed")
        if c is None or c[0] != sig:
            c = (sig, self._pack_eval(dtype))
            self.__dict__["_packed"] = c
            self.__dict__["_packed_fresh"] = True     # written by kernels that may still be in flight
        else:
            self.__dict__["_packed_fresh"] = False    # written during an earlier forward: static for every kernel launched from now on
        return c[1]

    def _nhwc(self, x, ctx):
        raise NotImplementedError

    def forward(self, input):
        ctx = _Ctx(self._precision(), self.training)
        x = _ToNHWC.apply(input, self._in_ld or input.shape[1], ctx.dtype)
        y = self._nhwc(x, ctx)
        return _ToNCHW.apply(y, self._out_c)


_GRAD_SINK = None   # set by parallel.FlatAdamTrainer while it runs backward: gradients are then accumulated by the kernels
                    # straight into the parameters' pre-assigned .grad views (no zero-fill / permute / AccumulateGrad launches)


def _sink(p):
    # non-leaf "parameters" (the folded weights of a 32-channel block) get their gradient through autograd, never through the sink
    return p.grad if (_GRAD_SINK is not None and p.is_leaf and p.grad is not None) else None


def _sunk(*params):
    if _GRAD_SINK is not None:
        for p in params:
            if p.is_leaf and p.grad is not None:
                _GRAD_SINK.ready(p)


def _side(fn, keep, *params):
    """Run fn() (weight-gradient launches that accumulate into the sinks of `params`) on the trainer's side stream when there is one and every
    parameter has a sink -- nothing is then returned to autograd, so nothing downstream waits for these kernels except the optimiser."""
    q = ops.SIDE
    if q is None or ops.WGRAD_FUSED_REDUCE or any(_sink(p) is None for p in params):   # (a grid-barrier kernel must never leave the main stream)
        return fn()
    return q.run(fn, keep)


def _ret(p, g):
    """What autograd gets for parameter p: nothing if the kernel already accumulated into p.grad."""
    return None if _sink(p) is not None else g


def _bn_apply(bn, y_raw, stats, count, drop=None, res=None):
    """Train-mode BatchNorm (+ Dropout2d multiplier, + residual) + ReLU of y_raw from the fp64 batch sums its producer accumulated:
    (out, mean, invstd); updates bn's running statistics and step counter like nn.BatchNorm2d.forward."""
    out, mean, invstd = ops.bn_apply(y_raw, stats, count, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, BN_MOMENTUM, BN_EPS,
                                     drop=drop, res=res, relu=True)
    _tick(bn)
    return out, mean, invstd


def _tick(bn):
    if not bn.__dict__.get("_erf_nbt_managed", False):
        bn.num_batches_tracked += 1


_PREPACKED_FRESH = False  # True while FlatAdamTrainer runs a forward right after its batched weight re-pack

_FIXED_DROPOUT = None  # {non_bottleneck_1d module: [N,C] Dropout2d multipliers} installed by fixed_dropout()


@contextlib.contextmanager
def fixed_dropout(scales):
    """Parity runs: use the given Dropout2d multipliers (0 or 1/(1-p), [N,C] per block) instead of drawing them."""
    global _FIXED_DROPOUT
    prev, _FIXED_DROPOUT = _FIXED_DROPOUT, scales
    try:
        yield
    finally:
        _FIXED_DROPOUT = prev


def draw_dropout(model, n, device):
    """All Dropout2d multipliers of one training forward in ONE bernoulli call per distinct p (instead of four small kernels per block,
    each a serialisation point between programmatically dependent kernels): {block: [N, C] fp32 in {0, 1/(1-p)}}.  Same distribution as
    nn.Dropout2d (one draw per (sample, channel)); the order in which the generator is consumed differs from the reference's."""
    groups = {}
    for m in model.modules():
        if isinstance(m, non_bottleneck_1d) and m.use_dropout and m.dropout.p != 0:
            groups.setdefault(float(m.dropout.p), []).append(m)
    drawn = {}
    for p, mods in groups.items():
        total = sum(n * m.chann for m in mods)
        flat = F.dropout(torch.ones(total, device=device), p, True)
        off = 0
        for m in mods:
            drawn[m] = flat[off:off + n * m.chann].view(n, m.chann)
            off += n * m.chann
    return drawn


_WEIGHTS_EPOCH = 0  # bumped by code that updates parameters through aliases torch's version counters cannot see


def bump_weights_epoch():
    global _WEIGHTS_EPOCH
    _WEIGHTS_EPOCH += 1


def _sig(mod):
    return (_WEIGHTS_EPOCH,) + tuple((t.data_ptr(), t._version) for t in list(mod.parameters()) + list(mod.buffers()))



def _bn_fold(bn: nn.BatchNorm2d):
    return ops.bn_eval_affine(bn.weight.detach().float(), bn.bias.detach().float(), bn.running_mean, bn.running_var, bn.eps)


# Weight images of the im2col convs packed AHEAD of the layers, on a second stream, at the start of a captured training step (the ~17 pack
# launches of a step are tiny but sat on the critical path, the data-gradient ones in the middle of backward).  `epoch` ties the images to ONE step:
# prepack_end() bumps it, so nothing (an eval forward, the next step) can pick up images packed from older weights.
_PREPACK = {"epoch": 0, "stream": None, "joined": True}


def prepack_begin(convs, dtype):
    """convs: [(_TcConv, weight)].  Call while a CUDA graph is being captured, after the optimiser's weights are final for this step."""
    main = torch.cuda.current_stream()
    s = ops.side_stream(main.device)
    s.wait_stream(main)
    with torch.cuda.stream(s):
        for tc, w in convs:
            tc.prepack(w, dtype, _PREPACK["epoch"])
    _PREPACK["stream"], _PREPACK["joined"] = s, False


def prepack_end():
    if not _PREPACK["joined"] and _PREPACK["stream"] is not None:      # nobody used an image: still join the fork
        torch.cuda.current_stream().wait_stream(_PREPACK["stream"])
    _PREPACK["joined"] = True
    _PREPACK["epoch"] += 1


class _TcConv:
    """tcgen05 (software-im2col) execution of one conv layer: forward and data gradient, with cached plans."""

    def __init__(self, spec: ConvSpec):
        self.spec = spec
        self.dspec = ops.dgrad_spec(spec)
        self._plans = None
        self._pre = None    # (epoch, weight pointer, forward images, data-gradient images) of prepack()

    def prepack(self, w, dtype, epoch):
        fwd = ops.im2col_pack(w, self.spec, False, self.plans()[0]) if self.fwd_ok(dtype) else None
        bwd = ops.im2col_pack(w, self.spec, True, self.plans()[1]) if self.bwd_ok(dtype) else None
        self._pre = (epoch, w.data_ptr(), fwd, bwd)

    def _prepacked(self, w, idx):
        pre = self._pre
        if pre is None or pre[0] != _PREPACK["epoch"] or pre[1] != w.data_ptr() or pre[idx] is None or _PREPACK["stream"] is None:
            return None
        if not _PREPACK["joined"]:      # first consumer of the step: the capturing stream waits for the packing stream once
            torch.cuda.current_stream().wait_stream(_PREPACK["stream"])
            _PREPACK["joined"] = True
        return pre[idx]

    def plans(self):
        if self._plans is None:
            self._plans = (ops.im2col_plan(self.spec), ops.im2col_plan(self.dspec))
        return self._plans

    def fwd_ok(self, dtype):
        return ops.im2col_supported(self.spec, dtype)

    def bwd_ok(self, dtype):
        return ops.im2col_supported(self.dspec, dtype)

    def pack_fwd(self, w):
        pre = self._prepacked(w, 2)
        return pre if pre is not None else ops.im2col_pack(w, self.spec, False, self.plans()[0])

    def forward(self, x, imgs, bias, relu=False, out=None, out_coff=0):
        return ops.conv_im2col_tc(x, imgs, self.plans()[0], self.spec, bias, relu=relu, out=out, out_coff=out_coff)

    def nchw_ok(self, dtype):
        return ops.USE_TC and ops.FUSE_NCHW_HEAD and ops.im2col_nchw_supported(self.spec, dtype)

    def nchw_plan(self):
        if self.__dict__.get("_nchw_plan") is None:
            self._nchw_plan = ops.im2col_plan_nchw(self.spec)
        return self._nchw_plan

    def forward_nchw(self, x, imgs, bias, relu=False):
        return ops.conv_im2col_tc_nchw(x, imgs, self.nchw_plan(), self.spec, bias, relu=relu)

    def dgrad(self, dy, w, out_hw):
        plan = self.plans()[1]
        imgs = self._prepacked(w, 3)
        if imgs is None:
            imgs = ops.im2col_pack(w, self.spec, True, plan)
        return ops.conv_im2col_tc(dy, imgs, plan, self.dspec, out_hw=out_hw)

    def wgrad(self, x, dy, wparam, bparam, bias_is_zero, bias_done=False):
        """(dw, db) for autograd, accumulated in place when a gradient sink is active.  bias_is_zero: the conv feeds a
        train-mode BatchNorm, so its bias gradient is exactly zero (see non_bottleneck_1d._bwd)."""
        return _side(lambda: self._wgrad_now(x, dy, wparam, bparam, bias_is_zero, bias_done), (x, dy), wparam, bparam)

    def _wgrad_now(self, x, dy, wparam, bparam, bias_is_zero, bias_done=False):
        """bias_done: the bias gradient was already accumulated into its sink by the producer of dy (the fused CE kernel)."""
        if ops.wgrad_im2col_supported(self.spec, x.dtype):
            dw = ops.conv_wgrad_im2col_tc(x, dy, self.spec, dw_out=_sink(wparam))
            bdst = _sink(bparam)
            if bias_done and bdst is not None:
                db = bdst
            elif bias_is_zero:
                db = None if bdst is not None else torch.zeros(self.spec.cout, device=x.device, dtype=torch.float32)
            else:
                db = bdst if bdst is not None else torch.zeros(self.spec.cout, device=x.device, dtype=torch.float32)
                c8 = ops.round_up(self.spec.cout, 8)
                if c8 != self.spec.cout and dy.shape[-1] >= c8:   # vectorised kernel over the padded channels, then keep the real ones
                    tmp = torch.zeros(c8, device=x.device, dtype=torch.float32)
                    ops.channel_sum_slice(dy, c8, tmp)
                    db += tmp[:self.spec.cout]
                else:
                    ops.channel_sum_slice(dy, self.spec.cout, db)
            return dw, db
        return ops.conv_wgrad(x, dy, self.spec, dw_out=_sink(wparam), db_out=_sink(bparam))


def _run(block, x, ctx):
    """Dispatch a block on an NHWC tensor: autograd path in training, folded kernels otherwise."""
    if ctx.training:
        return _BlockFn.apply(block, ctx, x, *block._train_params())
    with torch.no_grad():
        return block._fwd_eval(x, ctx)


# ---------------------------------------------------------------------------------------------------------------
class DownsamplerBlock(_B200Module):
    """eval/erfnet.py:12-23 -- cat([conv3x3 s2 p1 (ninput -> noutput-ninput), maxpool2x2]) -> BN(eps 1e-3) -> ReLU.
    One output tensor, two kernels writing its two channel slices."""

    def __init__(self, ninput, noutput, use_batchnorm=True):
        super().__init__()
        self.use_batchnorm = use_batchnorm
        self.conv = nn.Conv2d(ninput, noutput - ninput, (3, 3), stride=2, padding=1, bias=True)
        self.pool = nn.MaxPool2d(2, stride=2)
        self.bn = nn.BatchNorm2d(noutput, eps=BN_EPS)
        self.ninput, self.noutput = ninput, noutput
        self.spec = ConvSpec(ninput, noutput - ninput, 3, 3, 2, 1, 1)
        self.__dict__["_tc"] = _TcConv(self.spec)
        self._in_ld, self._out_c = ops.round_up(ninput, 4), noutput

    def _nhwc(self, x, ctx):
        return _run(self, x, ctx)

    def _check(self, x):
        if x.shape[1] % 2 or x.shape[2] % 2:
            # the reference fails in torch.cat here (conv gives ceil(H/2), pool floor(H/2)): SURVEY D8
            raise RuntimeError(f"DownsamplerBlock: H and W must be even, got {x.shape[1]}x{x.shape[2]}")

    def _pack_eval(self, dtype):
        w = self.conv.weight.detach().float()
        b = self.conv.bias.detach().float()
        cc = self.noutput - self.ninput
        if self.use_batchnorm:
            scale, shift = _bn_fold(self.bn)
            w = w * scale[:cc].view(-1, 1, 1, 1)
            b = b * scale[:cc] + shift[:cc]
            ps, psh = scale[cc:].contiguous(), shift[cc:].contiguous()
        else:
            ps = psh = None
        pk = dict(w=ops.pack_weight_fwd(w, self.spec, dtype), b=b.contiguous(), ps=ps, psh=psh)
        if self.ninput == 3 and self.noutput == 16:   # stem kernel operands: folded fp32 weight, affine of the pooled channels
            pk["stem_w"] = w.contiguous()
            if ps is not None:
                pk["stem_sc"] = torch.cat([torch.ones_like(b), ps]).contiguous()
                pk["stem_sh"] = torch.cat([torch.zeros_like(b), psh]).contiguous()
        if self._tc.fwd_ok(dtype):
            pk["tc"] = self._tc.pack_fwd(w)
        return pk

    def _fwd_eval(self, x, ctx):
        self._check(x)
        pk = self._cache(ctx.dtype)
        n, h, w, _ = x.shape
        if "stem_w" in pk and ops.stem_supported(x):
            return ops.stem_fwd(x, pk["stem_w"], pk["b"], pk.get("stem_sc"), pk.get("stem_sh"), relu=True)
        y = torch.empty(n, h // 2, w // 2, self.noutput, device=x.device, dtype=x.dtype)
        pool = lambda: ops.maxpool_fwd(x, self.ninput, y, self.noutput - self.ninput, pk["ps"], pk["psh"], relu=True)
        if "tc" in pk and ops.USE_TC:   # the two halves write disjoint channel slices: side by side under graph replay
            ops.fork_join(pool, lambda: self._tc.forward(x, pk["tc"], pk["b"], relu=True, out=y))
        else:
            ops.conv_fwd(x, pk["w"], self.spec, pk["b"], relu=True, out=y)
            pool()
        return y

    def _fwd_eval_nchw(self, input, ctx):
        """Eval-mode stem straight from the NCHW fp32 image (caller checked ops.stem_nchw_supported and that this is the 3 -> 16 block)."""
        pk = self._cache(ctx.dtype)
        return ops.stem_fwd_nchw(input, pk["stem_w"], pk["b"], pk.get("stem_sc"), pk.get("stem_sh"), True, ctx.dtype)

    def _train_params(self):
        return (self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)

    def _fwd_train(self, x, ctx):
        self._check(x)
        n, h, w, _ = x.shape
        cc = self.noutput - self.ninput
        stem = self.ninput == 3 and self.noutput == 16 and ops.stem_supported(x)
        if stem and not self.use_batchnorm:
            y = ops.stem_fwd(x, self.conv.weight.detach(), self.conv.bias.detach(), relu=True)
            return y, (x, None, y, None, None)
        if stem:
            stats = ops.bn_stats_buf(self.noutput, x.device)
            y_raw = ops.stem_fwd(x, self.conv.weight.detach(), self.conv.bias.detach(), stats=stats)
            y, mean, invstd = _bn_apply(self.bn, y_raw, stats, n * (h // 2) * (w // 2))
            return y, (x, y_raw, y, mean, invstd)
        y_raw = torch.empty(n, h // 2, w // 2, self.noutput, device=x.device, dtype=x.dtype)
        if not self.use_batchnorm:  # train/erfnet_blocks.py:24-29: relu(cat(conv, pool)), both ReLUs in the producing kernels
            if self._tc.fwd_ok(x.dtype):
                img, bias = self._tc.pack_fwd(self.conv.weight), self.conv.bias.detach().float()
                ops.fork_join(lambda: ops.maxpool_fwd(x, self.ninput, y_raw, cc, relu=True), lambda: self._tc.forward(x, img, bias, relu=True, out=y_raw))
            else:
                ops.conv_fwd(x, ops.pack_weight_fwd(self.conv.weight.detach(), self.spec, x.dtype), self.spec, self.conv.bias.detach().float(),
                             relu=True, out=y_raw)
                ops.maxpool_fwd(x, self.ninput, y_raw, cc, relu=True)
            return y_raw, (x, None, y_raw, None, None)
        stats = ops.bn_stats_buf(self.noutput, x.device)
        if self._tc.fwd_ok(x.dtype):
            img, bias = self._tc.pack_fwd(self.conv.weight), self.conv.bias.detach().float()
            ops.fork_join(lambda: ops.maxpool_fwd(x, self.ninput, y_raw, cc), lambda: self._tc.forward(x, img, bias, out=y_raw))
            ops.channel_stats(y_raw, stats)
        else:
            ops.conv_fwd(x, ops.pack_weight_fwd(self.conv.weight.detach(), self.spec, x.dtype), self.spec, self.conv.bias.detach().float(),
                         out=y_raw, stats=stats)
            ops.maxpool_fwd(x, self.ninput, y_raw, cc, stats=stats)
        y, mean, invstd = _bn_apply(self.bn, y_raw, stats, n * (h // 2) * (w // 2))
        return y, (x, y_raw, y, mean, invstd)

    def _bwd(self, dy, saved, need_dx):
        x, y_raw, y, mean, invstd = saved
        cw, cb, gw, gb = self._train_params()
        if self.use_batchnorm:
            dyr, dgamma, dbeta = ops.bn_bwd(dy, y_raw, gw.detach(), mean, invstd, mask=y, dgamma_out=_sink(gw), dbeta_out=_sink(gb), beta_for_mask=gb.detach())
        else:
            dyr, dgamma, dbeta = ops.relu_bwd(dy, y), None, None
        if self.ninput == 3 and self.noutput == 16 and ops.stem_supported(x) and not (ops.STEM_WGRAD_TC and ops.wgrad_im2col_supported(self.spec, x.dtype)):
            dw, db = _side(lambda: ops.stem_wgrad(x, dyr, dw_out=_sink(cw), db_out=_sink(cb), want_bias=not self.use_batchnorm), (x, dyr), cw, cb)
            if db is None and _sink(cb) is None:   # conv bias feeding a train-mode BN: gradient exactly zero
                db = torch.zeros(self.noutput - self.ninput, device=x.device, dtype=torch.float32)
        else:
            dw, db = self._tc.wgrad(x, dyr, cw, cb, bias_is_zero=self.use_batchnorm)
        dx = None
        if need_dx:
            n, h, w, _ = x.shape
            if self._tc.bwd_ok(x.dtype):
                dx = self._tc.dgrad(dyr, cw, (h, w))
            else:
                dx = ops.conv_fwd(dyr, ops.pack_weight_dgrad(cw.detach(), self.spec, x.dtype), ops.dgrad_spec(self.spec), out_hw=(h, w))
            ops.maxpool_bwd(x, self.ninput, dyr, self.noutput - self.ninput, dx)
        grads = (_ret(cw, dw), _ret(cb, db), _ret(gw, dgamma), _ret(gb, dbeta))
        _sunk(cw, cb, gw, gb)
        return dx, grads


# ---------------------------------------------------------------------------------------------------------------
class non_bottleneck_1d(_B200Module):
    """eval/erfnet.py:26-61 -- 3x1 -> ReLU -> 1x3 -> BN -> ReLU -> dilated 3x1 -> ReLU -> dilated 1x3 -> BN ->
    Dropout2d -> + input -> ReLU."""

    def __init__(self, chann, dropprob, dilated, use_dropout=True, use_batchnorm=True):
        super().__init__()
        self.use_dropout, self.use_batchnorm = use_dropout, use_batchnorm
        self.conv3x1_1 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1, 0), bias=True)
        self.conv1x3_1 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1), bias=True)
        self.bn1 = nn.BatchNorm2d(chann, eps=BN_EPS)
        self.conv3x1_2 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1 * dilated, 0), bias=True, dilation=(dilated, 1))
        self.conv1x3_2 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1 * dilated), bias=True, dilation=(1, dilated))
        self.bn2 = nn.BatchNorm2d(chann, eps=BN_EPS)
        self.dropout = nn.Dropout2d(dropprob)
        self.chann, self.dilated = chann, dilated
        d = dilated
        self.specs = (ConvSpec(chann, chann, 3, 1, 1, 1, 0), ConvSpec(chann, chann, 1, 3, 1, 0, 1),
                      ConvSpec(chann, chann, 3, 1, 1, d, 0, d, 1), ConvSpec(chann, chann, 1, 3, 1, 0, d, 1, d))
        self._out_c = chann

    def _convs(self):
        ov = self.__dict__.get("_conv_override")   # the 64-channel twin of a folded 32-channel block: (weight, bias) holders derived per step
        return ov if ov is not None else (self.conv3x1_1, self.conv1x3_1, self.conv3x1_2, self.conv1x3_2)

    def _nhwc(self, x, ctx):
        if self.chann == 32 and self._fold_ok(x, ctx):
            return self._nhwc_folded(x, ctx)
        return _run(self, x, ctx)

    # ---- 32-channel blocks (the fork's LadderBlock(16), train/erfnet_blocks.py:144) on the 64-channel tcgen05 kernels -----------------
    # [N,H,W,32] is the same memory as [N,H,W/2,64]: a "folded" pixel holds two neighbours (sub-pixel 0 = even column, 1 = odd column).
    # A 3x1 conv of the block is then a 3x1 conv over folded pixels with block-diagonal weights diag(W_t, W_t); a 1x3 conv (dilation 1) is
    # a 1x3 conv over folded pixels whose taps are  [-1]: out.sub0 <- in.sub1 * W_0;  [0]: [[W_1, W_2], [W_0, W_1]];  [+1]: out.sub1 <- in.sub0 * W_2.
    # Half of those MMAs multiply zeros, which is free next to running the block on the CUDA cores.  The folded weights are built with
    # differentiable torch ops, so autograd maps the 64-channel weight gradients back onto the 32-channel parameters.
    def _fold_ok(self, x, ctx):
        if self.use_batchnorm or self.dilated != 1 or not ops.USE_TC or not ops.FOLD_32 or x.shape[2] % 2:
            return False
        if ctx.training and self.use_dropout and self.dropout.p != 0:
            return False                      # Dropout2d masks are per original channel; not folded
        return ops.conv1d_tc_supported(64, x.shape[2] // 2, 1, x.dtype)

    def _folded_weights(self):
        """[(w64, b64)] x 4 from the 32-channel parameters (no autograd: _FoldAdapter maps the gradients back)."""
        out = []
        with torch.no_grad():
            for i, conv in enumerate((self.conv3x1_1, self.conv1x3_1, self.conv3x1_2, self.conv1x3_2)):
                w = conv.weight.detach()
                z = w.new_zeros(32, 32)
                blk = lambda a, b, c, d: torch.cat([torch.cat([a, b], 1), torch.cat([c, d], 1)], 0)     # [[a, b], [c, d]]: rows = out, cols = in
                if i % 2 == 0:
                    w64 = torch.stack([blk(w[:, :, t, 0], z, z, w[:, :, t, 0]) for t in range(3)], 2).unsqueeze(3)          # [64,64,3,1]
                else:
                    w0, w1, w2 = w[:, :, 0, 0], w[:, :, 0, 1], w[:, :, 0, 2]
                    w64 = torch.stack([blk(z, w0, z, z), blk(w1, w2, w0, w1), blk(z, z, w2, z)], 2).unsqueeze(2)            # [64,64,1,3]
                out.append((w64.contiguous(), torch.cat([conv.bias.detach(), conv.bias.detach()])))
        return out

    @staticmethod
    def _unfold_grads(i, dw64, db64):
        """Transpose of the folding map: gradients of the 64-channel twin -> gradients of conv i of the 32-channel block."""
        a, b = slice(0, 32), slice(32, 64)
        if i % 2 == 0:
            dw = (dw64[a, a] + dw64[b, b])                                                        # [32,32,3,1]
        else:
            g = dw64[:, :, 0, :]                                                                  # [64,64,3]
            dw = torch.stack([g[a, b, 0] + g[b, a, 1], g[a, a, 1] + g[b, b, 1], g[a, b, 1] + g[b, a, 2]], 2).unsqueeze(2)   # [32,32,1,3]
        return dw, db64[a] + db64[b]

    def _twin(self, device):
        tw = self.__dict__.get("_twin64")
        if tw is None or tw.conv3x1_1.weight.device != device:
            tw = non_bottleneck_1d(64, 0.0, 1, use_dropout=False, use_batchnorm=False).to(device)
            tw.precision = None
            self.__dict__["_twin64"] = tw     # not a registered sub-module: state_dict() / parameters() stay the reference's
        return tw

    def _nhwc_folded(self, x, ctx):
        n, h, w, _ = x.shape
        tw = self._twin(x.device)
        xf = x.view(n, h, w // 2, 64)
        if ctx.training:
            yf = _BlockFn.apply(_FoldAdapter(self, tw), ctx, xf, *self._train_params())
        else:
            sig = _sig(self)
            if self.__dict__.get("_twin_sig") != sig:
                with torch.no_grad():
                    for conv, (a, b) in zip((tw.conv3x1_1, tw.conv1x3_1, tw.conv3x1_2, tw.conv1x3_2), self._folded_weights()):
                        conv.weight.copy_(a)
                        conv.bias.copy_(b)
                self.__dict__["_twin_sig"] = sig
            tw.__dict__["_conv_override"] = None
            with torch.no_grad():
                yf = tw._fwd_eval(xf, ctx)
        return yf.view(n, h, w, 32)

    def _pack_eval(self, dtype):
        ws = [c.weight.detach().float() for c in self._convs()]
        bs = [c.bias.detach().float() for c in self._convs()]
        if self.use_batchnorm:
            for i, bn in ((1, self.bn1), (3, self.bn2)):
                scale, shift = _bn_fold(bn)
                ws[i] = ws[i] * scale.view(-1, 1, 1, 1)
                bs[i] = bs[i] * scale + shift
        pk = dict(w=[ops.pack_weight_fwd(w, s, dtype) for w, s in zip(ws, self.specs)], b=[b.contiguous() for b in bs])
        if dtype == torch.bfloat16 and self.chann in (16, 64, 128) and not ops.EMULATED:
            pk["tc"] = [ops.pack_conv1d_tc(w) for w in ws]  # tcgen05 operand images (BN folded)
        return pk

    def _fwd_eval(self, x, ctx):
        pk = self._cache(ctx.dtype)
        w, b, s = pk["w"], pk["b"], self.specs
        d = self.dilated
        if "tc" in pk and ops.conv1d_tc_supported(self.chann, x.shape[2], d, x.dtype) and ops.USE_TC:
            tc = pk["tc"]
            st = ops.STATIC_WEIGHT_PREFETCH and not self.__dict__.get("_packed_fresh", True)   # cached images: their copy may start before pdl_wait()
            if ops.conv1d_pair_tc_supported(self.chann, x.shape[2], d, x.dtype) and ops.conv1d_pair_tc_supported(self.chann, x.shape[2], 1, x.dtype):
                t = ops.conv1d_pair_tc(x, tc[0], b[0], tc[1], b[1], 1, relu=True, static_w=st)            # 3x1 -> ReLU -> 1x3 (+BN1 folded) -> ReLU
                return ops.conv1d_pair_tc(t, tc[2], b[2], tc[3], b[3], d, relu=True, res=x, static_w=st)  # dilated pair (+BN2 folded) + x -> ReLU
            t = ops.conv1d_tc(x, tc[0], b[0], False, 1, relu=True, static_w=st)
            t = ops.conv1d_tc(t, tc[1], b[1], True, 1, relu=True, static_w=st)
            t = ops.conv1d_tc(t, tc[2], b[2], False, d, relu=True, static_w=st)
            return ops.conv1d_tc(t, tc[3], b[3], True, d, relu=True, res=x, static_w=st)
        t = ops.conv_fwd(x, w[0], s[0], b[0], relu=True)
        t = ops.conv_fwd(t, w[1], s[1], b[1], relu=True)
        t = ops.conv_fwd(t, w[2], s[2], b[2], relu=True)
        return ops.conv_fwd(t, w[3], s[3], b[3], relu=True, add=x)

    def _train_params(self):
        c = self._convs()
        return (c[0].weight, c[0].bias, c[1].weight, c[1].bias, self.bn1.weight, self.bn1.bias,
                c[2].weight, c[2].bias, c[3].weight, c[3].bias, self.bn2.weight, self.bn2.bias)

    def _drop_scale(self, x, ctx):
        """Dropout2d multiplier [N,C] (0 or 1/(1-p)): same RNG consumption as nn.Dropout2d on [N,C,H,W] (noise [N,C,1,1])."""
        p = self.dropout.p
        if not (self.use_dropout and p != 0):  # eval/erfnet.py:58
            return None
        fixed = ctx.drop_scales if ctx.drop_scales is not None else _FIXED_DROPOUT
        if fixed is not None and self in fixed:
            return fixed[self].to(x.device, torch.float32).reshape(x.shape[0], self.chann).contiguous()
        if ctx.drawn is not None and self in ctx.drawn:      # drawn for the whole net in one call (draw_dropout)
            return ctx.drawn[self]
        ones = torch.ones(x.shape[0], self.chann, 1, 1, device=x.device)
        return F.dropout2d(ones, p, True).reshape(x.shape[0], self.chann).contiguous()

    def _use_tc(self, x):
        return (self.chann in (16, 64, 128) and ops.USE_TC and ops.conv1d_tc_supported(self.chann, x.shape[2], self.dilated, x.dtype)
                and ops.conv1d_tc_supported(self.chann, x.shape[2], 1, x.dtype))

    def _conv(self, i, x, w, bias, tc, relu=False, stats=None, mask=None, add=None, spec=None, colsum=None, static_w=False):
        """conv i (0..3) of the block (or its data gradient when spec is a dgrad spec) on the tcgen05 or the generic kernel.
        colsum: fp32 [C] that receives the per-channel sum of the result (= bias gradient of the conv that produced the
        tensor this gradient belongs to)."""
        if tc:
            return ops.conv1d_tc(x, w, bias, horizontal=bool(i & 1), dil=1 if i < 2 else self.dilated, relu=relu, res=add, mask=mask,
                                 stats=stats, colsum=colsum, static_w=static_w and ops.STATIC_WEIGHT_PREFETCH)
        y = ops.conv_fwd(x, w, spec or self.specs[i], bias, relu=relu, stats=stats, mask=mask, add=add, out_hw=x.shape[1:3])
        if colsum is not None:
            ops.channel_sum(y, colsum)
        return y

    def _fwd_train(self, x, ctx):
        n, h, w, c = x.shape
        s, dtp = self.specs, x.dtype
        cv = self._convs()
        tc = self._use_tc(x)
        if tc:
            # forward + data-gradient images: pre-packed for this step by the trainer's batched launch, else one launch per conv
            pairs = self.__dict__.get("_prepacked") if (_PREPACKED_FRESH and self.__dict__.get("_conv_override") is None) else None
            sw = pairs is not None     # the trainer's images: complete before any kernel of this step starts (FlatAdamTrainer._begin_step)
            if pairs is None:
                pairs = [ops.pack_conv1d_tc_pair(m.weight) for m in cv]
            wp = [pr[0] for pr in pairs]
        else:
            pairs, sw = None, False
            wp = [ops.pack_weight_fwd(m.weight.detach(), sp, dtp) for m, sp in zip(cv, s)]
        bs = [m.bias.detach().float() for m in cv]
        if not self.use_batchnorm:
            # train/erfnet_blocks.py:55-73 without the BNs: conv+bias+ReLU chain, everything in the conv epilogues
            t1 = self._conv(0, x, wp[0], bs[0], tc, relu=True, static_w=sw)
            a2 = self._conv(1, t1, wp[1], bs[1], tc, relu=True, static_w=sw)
            t3 = self._conv(2, a2, wp[2], bs[2], tc, relu=True, static_w=sw)
            drop = self._drop_scale(x, ctx)
            if drop is None:
                out = self._conv(3, t3, wp[3], bs[3], tc, relu=True, add=x, static_w=sw)
            else:
                y4 = self._conv(3, t3, wp[3], bs[3], tc, static_w=sw)
                out = ops.affine_act(y4, drop, torch.zeros_like(drop), res=x, relu=True)
            return out, (x, t1, None, a2, t3, None, out, None, None, None, None, drop, (pairs, sw))
        count = n * h * w
        st = (ops.bn_stats_buf(c, x.device), ops.bn_stats_buf(c, x.device))
        t1 = self._conv(0, x, wp[0], bs[0], tc, relu=True, static_w=sw)
        y2 = self._conv(1, t1, wp[1], bs[1], tc, stats=st[0], static_w=sw)
        a2, mean1, invstd1 = _bn_apply(self.bn1, y2, st[0], count)
        t3 = self._conv(2, a2, wp[2], bs[2], tc, relu=True, static_w=sw)
        y4 = self._conv(3, t3, wp[3], bs[3], tc, stats=st[1], static_w=sw)
        drop = self._drop_scale(x, ctx)
        out, mean2, invstd2 = _bn_apply(self.bn2, y4, st[1], count, drop=drop, res=x)
        return out, (x, t1, y2, a2, t3, y4, out, mean1, invstd1, mean2, invstd2, drop, (pairs, sw))

    def _bwd(self, dout, saved, need_dx):
        x, t1, y2, a2, t3, y4, out, mean1, invstd1, mean2, invstd2, drop, (pairs, sw) = saved
        s, dtp = self.specs, x.dtype
        cv = self._convs()
        tc = self._use_tc(x)
        ds = [ops.dgrad_spec(sp) for sp in s]
        if tc:
            wd = [pr[1] for pr in pairs] if pairs is not None else [ops.pack_conv1d_tc(m.weight.detach(), dgrad=True) for m in cv]
        else:
            wd = [ops.pack_weight_dgrad(m.weight.detach(), sp, dtp) for m, sp in zip(cv, s)]
        P = self._train_params()  # w1,b1,w2,b2,g1,be1,w3,b3,w4,b4,g2,be2
        dev = x.device

        def bias_dst(p):
            d = _sink(p)
            return d if d is not None else torch.zeros(self.chann, device=dev, dtype=torch.float32)

        # The biases of conv1x3_1 / conv1x3_2 feed a train-mode BatchNorm: their gradient sum_p dy is EXACTLY zero
        # (sum_p xhat = 0), the reference only ever sees its rounding noise (~1e-9, far below Adam's eps = 1e-4) -> no kernel.
        if not self.use_batchnorm:
            return self._bwd_nobn(dout, saved, need_dx, tc, wd, ds, P, bias_dst, sw)
        db2 = db4 = None if (_sink(P[3]) is not None) else torch.zeros(self.chann, device=dev, dtype=torch.float32)
        db1, db3 = bias_dst(P[1]), bias_dst(P[7])
        dy4, dg2, dbe2, gres = ops.bn_bwd(dout, y4, P[10].detach(), mean2, invstd2, mask=out, drop=drop, dgamma_out=_sink(P[10]),
                                          dbeta_out=_sink(P[11]), want_masked=True)  # gres = dout * (out > 0): the residual path's gradient
        grp = self._wgrad_group(x, tc, P)
        dw4, _ = self._wgrad(3, t3, dy4, tc, P[8], grp)
        dt3 = self._conv(3, dy4, wd[3], None, tc, mask=t3, spec=ds[3], colsum=db3, static_w=sw)     # grad wrt conv3x1_2's output
        dw3, _ = self._wgrad(2, a2, dt3, tc, P[6], grp)
        da2 = self._conv(2, dt3, wd[2], None, tc, spec=ds[2], static_w=sw)
        dy2, dg1, dbe1 = ops.bn_bwd(da2, y2, P[4].detach(), mean1, invstd1, mask=a2, dgamma_out=_sink(P[4]), dbeta_out=_sink(P[5]),
                                     beta_for_mask=P[5].detach())
        dw2, _ = self._wgrad(1, t1, dy2, tc, P[2], grp)
        dt1 = self._conv(1, dy2, wd[1], None, tc, mask=t1, spec=ds[1], colsum=db1, static_w=sw)     # grad wrt conv3x1_1's output
        dw1, _ = self._wgrad(0, x, dt1, tc, P[0], grp)
        dx = None
        if need_dx:
            dx = self._conv(0, dt1, wd[0], None, tc, add=gres, spec=ds[0], static_w=sw)
        if grp:
            ops.conv1d_wgrad_tc_group(grp)
        grads = tuple(_ret(p, g) for p, g in zip(P, (dw1, db1, dw2, db2, dg1, dbe1, dw3, db3, dw4, db4, dg2, dbe2)))
        _sunk(*P)
        return dx, grads

    def _bwd_nobn(self, dout, saved, need_dx, tc, wd, ds, P, bias_dst, sw=False):
        """Backward of the BN-free block: every ReLU mask and bias-gradient column sum rides in a data-gradient epilogue;
        the BN parameters get no gradient (they are unused, as in the reference)."""
        x, t1, _, a2, t3, _, out, _, _, _, _, drop, _ = saved
        db1, db2, db3, db4 = bias_dst(P[1]), bias_dst(P[3]), bias_dst(P[7]), bias_dst(P[9])
        gres = ops.relu_bwd(dout, out)                                                  # grad wrt (branch + x)
        dy4 = gres if drop is None else ops.affine_act(gres, drop, torch.zeros_like(drop), relu=False)
        ops.channel_sum(dy4, db4)
        grp = self._wgrad_group(x, tc, P)
        dw4, _ = self._wgrad(3, t3, dy4, tc, P[8], grp)
        dt3 = self._conv(3, dy4, wd[3], None, tc, mask=t3, spec=ds[3], colsum=db3, static_w=sw)
        dw3, _ = self._wgrad(2, a2, dt3, tc, P[6], grp)
        dy2 = self._conv(2, dt3, wd[2], None, tc, mask=a2, spec=ds[2], colsum=db2, static_w=sw)
        dw2, _ = self._wgrad(1, t1, dy2, tc, P[2], grp)
        dt1 = self._conv(1, dy2, wd[1], None, tc, mask=t1, spec=ds[1], colsum=db1, static_w=sw)
        dw1, _ = self._wgrad(0, x, dt1, tc, P[0], grp)
        dx = self._conv(0, dt1, wd[0], None, tc, add=gres, spec=ds[0], static_w=sw) if need_dx else None
        if grp:
            ops.conv1d_wgrad_tc_group(grp)
        grads = tuple(_ret(p, g) for p, g in zip(P, (dw1, db1, dw2, db2, None, None, dw3, db3, dw4, db4, None, None)))
        _sunk(*P)  # the unused BN parameters count as done (their gradient stays zero)
        return dx, grads

    def _wgrad(self, i, x, dy, tc, wparam, group=None):
        """Weight gradient only (the bias gradients come from the data-gradient epilogues, see _bwd).  group: a list that collects the block's
        weight-gradient jobs for ONE grouped launch at the end of its backward (only the optimiser waits for a weight gradient)."""
        if group is not None:
            group.append((x, dy, bool(i & 1), 1 if i < 2 else self.dilated, _sink(wparam)))
            return _sink(wparam), None
        return _side(lambda: self._wgrad_now(i, x, dy, tc, wparam), (x, dy), wparam)

    def _wgrad_group(self, x, tc, P):
        """[] when this block's four weight gradients can go out as one grouped launch (tensor-core path, every weight has a gradient sink), else
        None (one launch per conv, results returned to autograd)."""
        ok = (tc and ops.WGRAD_GROUP and not ops.EMULATED and ops.SIDE is None and ops.conv1d_wgrad_tc_supported(self.chann, x.shape[2], self.dilated, x.dtype)
              and all(_sink(P[i]) is not None for i in (0, 2, 6, 8)))
        return [] if ok else None

    def _wgrad_now(self, i, x, dy, tc, wparam):
        if tc and ops.conv1d_wgrad_tc_supported(self.chann, x.shape[2], self.dilated, x.dtype):
            return ops.conv1d_wgrad_tc(x, dy, horizontal=bool(i & 1), dil=1 if i < 2 else self.dilated, dw_out=_sink(wparam), want_bias=False)
        return ops.conv_wgrad(x, dy, self.specs[i], dw_out=_sink(wparam), want_bias=False)


# ---------------------------------------------------------------------------------------------------------------
class UpsamplerBlock(_B200Module):
    """eval/erfnet.py:99-108 -- ConvTranspose2d(3, stride 2, padding 1, output_padding 1) -> BN -> ReLU."""

    def __init__(self, ninput, noutput, use_batchnorm=True):
        super().__init__()
        self.use_batchnorm = use_batchnorm
        self.conv = nn.ConvTranspose2d(ninput, noutput, 3, stride=2, padding=1, output_padding=1, bias=True)
        self.bn = nn.BatchNorm2d(noutput, eps=BN_EPS)
        self.spec = ConvSpec(ninput, noutput, 3, 3, 2, 1, 1, transposed=True, out_pad=1)
        self.__dict__["_tc"] = _TcConv(self.spec)
        self._out_c = noutput

    def _nhwc(self, x, ctx):
        return _run(self, x, ctx)

    def _pack_eval(self, dtype):
        w, b = self.conv.weight.detach().float(), self.conv.bias.detach().float()
        if self.use_batchnorm:
            scale, shift = _bn_fold(self.bn)
            w = w * scale.view(1, -1, 1, 1)
            b = b * scale + shift
        pk = dict(w=ops.pack_weight_fwd(w, self.spec, dtype), b=b.contiguous())
        if self._tc.fwd_ok(dtype):
            pk["tc"] = self._tc.pack_fwd(w)
        return pk

    def _fwd_eval(self, x, ctx):
        pk = self._cache(ctx.dtype)
        if "tc" in pk and ops.USE_TC:
            return self._tc.forward(x, pk["tc"], pk["b"], relu=True)
        return ops.conv_fwd(x, pk["w"], self.spec, pk["b"], relu=True)

    def _train_params(self):
        return (self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)

    def _fwd_train(self, x, ctx):
        if not self.use_batchnorm:  # train/erfnet_blocks.py:84-88: relu(convT(x))
            if self._tc.fwd_ok(x.dtype):
                y = self._tc.forward(x, self._tc.pack_fwd(self.conv.weight), self.conv.bias.detach().float(), relu=True)
            else:
                y = ops.conv_fwd(x, ops.pack_weight_fwd(self.conv.weight.detach(), self.spec, x.dtype), self.spec, self.conv.bias.detach().float(),
                                 relu=True)
            return y, (x, None, y, None, None)
        stats = ops.bn_stats_buf(self.bn.num_features, x.device)
        if self._tc.fwd_ok(x.dtype):
            y_raw = self._tc.forward(x, self._tc.pack_fwd(self.conv.weight), self.conv.bias.detach().float())
            ops.channel_stats(y_raw, stats)
        else:
            y_raw = ops.conv_fwd(x, ops.pack_weight_fwd(self.conv.weight.detach(), self.spec, x.dtype), self.spec, self.conv.bias.detach().float(),
                                 stats=stats)
        count = y_raw.shape[0] * y_raw.shape[1] * y_raw.shape[2]
        y, mean, invstd = _bn_apply(self.bn, y_raw, stats, count)
        return y, (x, y_raw, y, mean, invstd)

    def _bwd(self, dy, saved, need_dx):
        x, y_raw, y, mean, invstd = saved
        cw, cb, gw, gb = self._train_params()
        if self.use_batchnorm:
            dyr, dgamma, dbeta = ops.bn_bwd(dy, y_raw, gw.detach(), mean, invstd, mask=y, dgamma_out=_sink(gw), dbeta_out=_sink(gb), beta_for_mask=gb.detach())
        else:
            dyr, dgamma, dbeta = ops.relu_bwd(dy, y), None, None
        dw, db = self._tc.wgrad(x, dyr, cw, cb, bias_is_zero=self.use_batchnorm)
        dx = None
        if need_dx:
            if self._tc.bwd_ok(x.dtype):
                dx = self._tc.dgrad(dyr, cw, x.shape[1:3])
            else:
                dx = ops.conv_fwd(dyr, ops.pack_weight_dgrad(cw.detach(), self.spec, x.dtype), ops.dgrad_spec(self.spec), out_hw=x.shape[1:3])
        grads = (_ret(cw, dw), _ret(cb, db), _ret(gw, dgamma), _ret(gb, dbeta))
        _sunk(cw, cb, gw, gb)
        return dx, grads


# ---------------------------------------------------------------------------------------------------------------
class _PlainConv:
    """A bare conv / transposed conv with bias (the two output_conv heads, eval/erfnet.py:85,124).  Built per call
    around the owner's nn.Conv2d / nn.ConvTranspose2d so that module replicas (DataParallel) stay independent."""

    def __init__(self, conv: nn.Module, spec: ConvSpec, relu: bool = False):
        self.spec = spec
        self._conv = conv
        self._tc = _TcConv(spec)
        self.out_ld = ops.round_up(spec.cout, 8)
        self.relu = relu   # conv + bias + ReLU (the 3x3 conv of the fork's LadderBlock, train/erfnet_blocks.py:150-154)

    def _out(self, x, tc=False):
        n, h, w, _ = x.shape
        ho, wo = self.spec.out_hw(h, w)
        if self.out_ld != self.spec.cout and not tc:   # the tcgen05 kernel writes the pad channels itself (zeros)
            return torch.zeros(n, ho, wo, self.out_ld, device=x.device, dtype=x.dtype)
        return torch.empty(n, ho, wo, self.out_ld, device=x.device, dtype=x.dtype)

    def _fwd_eval(self, x, ctx):
        sig = (ctx.dtype,) + _sig(self._conv)
        c = self._conv.__dict__.get("_packed")
        if c is None or c[0] != sig:
            tc = self._tc.pack_fwd(self._conv.weight) if self._tc.fwd_ok(ctx.dtype) else None
            c = (sig, (ops.pack_weight_fwd(self._conv.weight.detach().float(), self.spec, ctx.dtype), self._conv.bias.detach().float().contiguous(), tc))
            self._conv.__dict__["_packed"] = c
        w, b, tc = c[1]
        if tc is not None and ops.USE_TC:
            return self._tc.forward(x, tc, b, relu=self.relu, out=self._out(x, tc=True))
        return ops.conv_fwd(x, w, self.spec, b, relu=self.relu, out=self._out(x))

    def _fwd_eval_nchw(self, x, ctx):
        """Eval-mode forward whose epilogue writes the NCHW fp32 result directly (caller checked self._tc.nchw_ok)."""
        sig = (ctx.dtype,) + _sig(self._conv)
        c = self._conv.__dict__.get("_packed_nchw")
        if c is None or c[0] != sig:
            c = (sig, (ops.im2col_pack(self._conv.weight, self.spec, False, self._tc.nchw_plan()), self._conv.bias.detach().float().contiguous()))
            self._conv.__dict__["_packed_nchw"] = c
        imgs, b = c[1]
        return self._tc.forward_nchw(x, imgs, b, relu=self.relu)

    def _train_params(self):
        return (self._conv.weight, self._conv.bias)

    def _fwd_train(self, x, ctx):
        if self._tc.fwd_ok(x.dtype):
            y = self._tc.forward(x, self._tc.pack_fwd(self._conv.weight), self._conv.bias.detach().float(), relu=self.relu, out=self._out(x, tc=True))
        else:
            y = ops.conv_fwd(x, ops.pack_weight_fwd(self._conv.weight.detach(), self.spec, x.dtype), self.spec, self._conv.bias.detach().float(),
                             relu=self.relu, out=self._out(x))
        return y, (x, y if self.relu else None)

    def _bwd(self, dy, saved, need_dx):
        x, y = saved
        if self.relu:
            dy = ops.relu_bwd(dy, y)
        cw, cb = self._train_params()
        dw, db = self._tc.wgrad(x, dy, cw, cb, bias_is_zero=False, bias_done=self._conv.__dict__.pop("_erf_bias_grad_done", False))
        dx = None
        if need_dx:
            if self._tc.bwd_ok(x.dtype):
                dx = self._tc.dgrad(dy, cw, x.shape[1:3])
            elif ops.convT2x2_dgrad_tc_supported(self.spec, dy):
                dx = ops.convT2x2_dgrad_tc(dy, cw, self.spec)
            else:
                dx = ops.conv_fwd(dy, ops.pack_weight_dgrad(cw.detach(), self.spec, x.dtype), ops.dgrad_spec(self.spec), out_hw=x.shape[1:3])
        grads = (_ret(cw, dw), _ret(cb, db))
        _sunk(cw, cb)
        return dx, grads


# ---------------------------------------------------------------------------------------------------------------
class Encoder(_B200Module):
    """eval/erfnet.py:64-96."""

    def __init__(self, num_classes):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64))
        for x in range(0, 5):
            self.layers.append(non_bottleneck_1d(64, 0.1, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for x in range(0, 2):
            self.layers.append(non_bottleneck_1d(128, 0.1, 2))
            self.layers.append(non_bottleneck_1d(128, 0.1, 4))
            self.layers.append(non_bottleneck_1d(128, 0.1, 8))
            self.layers.append(non_bottleneck_1d(128, 0.1, 16))
        # only for encoder mode:
        self.output_conv = nn.Conv2d(128, num_classes, 1, stride=1, padding=0, bias=True)
        self.num_classes = num_classes
        self._in_ld = 4

    def _nhwc(self, x, ctx, predict=False, skip_stem=False):
        if not skip_stem:
            if x.shape[1] % 8 or x.shape[2] % 8:
                raise RuntimeError(f"ERFNet: H and W must be multiples of 8, got {x.shape[1]}x{x.shape[2]} (SURVEY D8)")
            y = _run(self.initial_block, x, ctx)
        else:
            y = x
        for layer in self.layers:
            y = _run(layer, y, ctx)
        if predict:
            y = _run(_PlainConv(self.output_conv, ConvSpec(128, self.num_classes, 1, 1)), y, ctx)
        return y

    def _from_input(self, input, ctx, predict=False):
        """Public-boundary input (NCHW fp32, or uint8 HWC) -> encoder output (NHWC).  In inference the stem reads an NCHW fp32 image itself."""
        ib = self.initial_block
        if (not ctx.training and input.dtype == torch.float32 and ops.stem_nchw_supported(input) and ib.ninput == 3 and ib.noutput == 16
                and input.shape[3] <= 4096):
            if input.shape[2] % 8 or input.shape[3] % 8:
                raise RuntimeError(f"ERFNet: H and W must be multiples of 8, got {input.shape[2]}x{input.shape[3]} (SURVEY D8)")
            with torch.no_grad():
                y = ib._fwd_eval_nchw(input, ctx)
            return self._nhwc(y, ctx, predict, skip_stem=True)
        return self._nhwc(input_to_nhwc(input, 4, ctx.dtype), ctx, predict)

    def forward(self, input, predict=False):
        ctx = _Ctx(self._precision(), self.training)
        y = self._from_input(input, ctx, predict)
        return _ToNCHW.apply(y, self.num_classes if predict else 128)


class Decoder(_B200Module):
    """eval/erfnet.py:110-134."""

    def __init__(self, num_classes):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(UpsamplerBlock(128, 64))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(UpsamplerBlock(64, 16))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.output_conv = nn.ConvTranspose2d(16, num_classes, 2, stride=2, padding=0, output_padding=0, bias=True)
        self.num_classes = num_classes
        self._out_c = num_classes

    def _nhwc(self, x, ctx):
        y = x
        for layer in self.layers:
            y = _run(layer, y, ctx)
        return _run(_PlainConv(self.output_conv, ConvSpec(16, self.num_classes, 2, 2, 2, transposed=True)), y, ctx)


class ERFNet(_B200Module):
    """eval/erfnet.py:137-152.  ``ERFNet(num_classes, encoder=None).forward(input, only_encode=False)``."""

    def __init__(self, num_classes, encoder=None):
        super().__init__()
        if encoder is None:
            self.encoder = Encoder(num_classes)
        else:
            self.encoder = encoder
        self.decoder = Decoder(num_classes)
        self.num_classes = num_classes

    def forward_nhwc(self, input, only_encode=False, drop_scales=None):
        """NCHW fp32 in, internal NHWC logits out ([N,H,W,round8(C)], activation dtype) -- what the fused loss /
        argmax / histogram kernels consume without the NCHW fp32 round trip."""
        ctx = _Ctx(self._precision(), self.training, drop_scales)
        if self.training and drop_scales is None and _FIXED_DROPOUT is None and input.is_cuda and not ops.EMULATED:
            ctx.drawn = draw_dropout(self, input.shape[0], input.device)
        if only_encode:
            return self._encode(input, ctx, predict=True)
        return self.decoder._nhwc(self._encode(input, ctx), ctx)

    def _encode(self, input, ctx, predict=False):
        """Encoder pass from the public-boundary input; also serves a foreign encoder handed to the constructor, e.g. the ImageNet
        pre-trained ``features.encoder`` (train/main.py:857-867), which has no prediction conv."""
        enc = self.encoder
        if hasattr(enc, "_from_input"):
            return enc._from_input(input, ctx, predict)
        x = input_to_nhwc(input, 4, ctx.dtype)
        return enc._nhwc(x, ctx, True) if predict else enc._nhwc(x, ctx)

    def forward(self, input, only_encode=False):
        if not self.training and not only_encode:
            # inference: the head's epilogue writes the NCHW fp32 logits itself when it runs on the tensor cores
            ctx = _Ctx(self._precision(), False)
            head = _PlainConv(self.decoder.output_conv, ConvSpec(16, self.num_classes, 2, 2, 2, transposed=True))
            if head._tc.nchw_ok(ctx.dtype) and input.is_cuda:
                with torch.no_grad():
                    y = self._encode(input, ctx)
                    for layer in self.decoder.layers:
                        y = _run(layer, y, ctx)
                    return head._fwd_eval_nchw(y, ctx)
        y = self.forward_nhwc(input, only_encode)
        return _ToNCHW.apply(y, self.num_classes)


Net = ERFNet  # train/main.py instantiates ``model_file.Net``; the Cityscapes graph is the ERFNet above
