"""One-process-per-GPU data-parallel training of ERFNet (replaces nn.DataParallel).

Reference behaviour being reproduced (train/main.py:297-392 under nn.DataParallel):
  * the loss is taken over the GATHERED global batch -> here every rank normalises its cross-entropy by the
    all-reduced sum of valid-pixel weights, and gradients are SUM-all-reduced (not averaged);
  * BatchNorm statistics are per replica (no SyncBN); running stats survive from replica 0;
  * Adam(lr 1e-4, betas (.9,.999), eps 1e-4, weight_decay 1e-6) with L2 in the gradient (train/main.py:43-46,213).

B200 design: parameters, gradients and Adam moments live in ONE flat fp32 buffer each (2.07 M values); gradients
are all-reduced in 3 contiguous buckets in reverse graph order (decoder, 128-ch encoder, 64-ch encoder + stem),
each launched on a side stream from a post-accumulate-grad hook as soon as its last gradient is written, so NCCL
over NVLink overlaps the rest of backward; the optimizer is one fused kernel over the flat buffers.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import erfnet as _erfnet
from . import ops
from .evaluation_functions import CrossEntropyLoss2d


def _bucket_of(name: str) -> int:
    if name.startswith("decoder."):
        return 0
    if name.startswith("encoder.layers."):
        idx = int(name.split(".")[2])
        return 1 if idx >= 6 else 2
    return 2  # initial_block, encoder.output_conv


class FlatAdamTrainer:
    """Fused training step: forward (NHWC logits) -> CE fwd+bwd kernel -> backward -> bucketed all-reduce -> flat Adam."""

    def __init__(self, model, weight=None, ignore_index=-1, lr=1e-4, betas=(0.9, 0.999), eps=1e-4, weight_decay=1e-6,
                 process_group=None, num_classes=None, loss_fn=None, sgd_momentum=None):
        """loss_fn: None = the fused Cityscapes path (ERFNet.forward_nhwc + weighted CE kernel).  Otherwise any of the nets with any
        criterion: ``loss_fn(model(images), *targets)`` -> scalar loss (e.g. ``lambda out, t: L1LossMasked()(out, t)`` for the fork's
        regression nets, train/main.py:174-184).  At world_size > 1 this package's criteria normalise by the GLOBAL batch (their (sum, count)
        accumulators are all-reduced, ops.LOSS_REDUCE), which with SUM-all-reduced gradients is exactly the reference's loss over the outputs
        gathered by nn.DataParallel; a foreign criterion's per-rank means are averaged instead."""
        self.model = model
        self.loss_fn = loss_fn
        self.sgd_momentum = sgd_momentum   # not None: torch.optim.SGD(lr, momentum, weight_decay) semantics (imagenet/main.py:83-85) instead of Adam
        self.num_classes = num_classes or getattr(model, "num_classes", None)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self._hyper_dev = None      # fp32 [2] = {lr, weight_decay} on the device once a CUDA graph holds the optimiser launch (set_lr keeps it current)
        self._lr, self._wd = float(lr), float(weight_decay)
        self.betas, self.eps = betas, eps
        self.criterion = CrossEntropyLoss2d(weight, ignore_index)
        self.step_count = 0
        # The trainer takes ownership of the parameters: every nn.Parameter is replaced by a FRESH one (a view of the flat buffer below).
        # A grad-enabled forward that ran before this point leaves AccumulateGrad nodes bound to the stream it ran on (normally the legacy
        # default stream); the autograd engine would then synchronise that stream with the capturing stream at the end of a captured
        # backward ("dependency created on uncaptured work in another stream").  Fresh parameters have no such history.
        fresh = {}
        for mod in model.modules():
            for pname, par in list(mod._parameters.items()):
                if par is not None:
                    if id(par) not in fresh:
                        fresh[id(par)] = torch.nn.Parameter(par.detach(), requires_grad=par.requires_grad)
                    mod._parameters[pname] = fresh[id(par)]
        named = list(model.named_parameters())
        # flat layout ordered by bucket so that each bucket is one contiguous slice
        order = sorted(range(len(named)), key=lambda i: (_bucket_of(named[i][0]), i))
        # every parameter starts on a 16-byte boundary of the flat buffers (the weight-gradient kernels add into their slice with TMA bulk reductions,
        # the optimiser kernels use 16-byte accesses); the few padding values stay zero
        total = sum((p.numel() + 3) // 4 * 4 for _, p in named)
        dev = named[0][1].device
        self.criterion = self.criterion.to(dev)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.bucket_slices: List[List[int]] = [[None, None] for _ in range(3)]
        self._bucket_params = [0, 0, 0]
        self._param_bucket = {}
        off = 0
        for i in order:
            name, p = named[i]
            n = p.numel()
            b = _bucket_of(name)
            self.flat_p[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
            p.grad = self.flat_g[off:off + n].view_as(p)
            sl = self.bucket_slices[b]
            sl[0] = off if sl[0] is None else sl[0]
            sl[1] = off + n
            self._bucket_params[b] += 1
            self._param_bucket[p] = b
            off += (n + 3) // 4 * 4
        self._params = [p for _, p in named]                         # model.parameters() order = torch.optim's parameter indices
        self._offsets = {}
        for i in order:
            p = named[i][1]
            self._offsets[p] = p.data.storage_offset() - self.flat_p.storage_offset()
        # all BatchNorm step counters in one int64 tensor: one increment per step instead of 39
        unused = {c for mod in model.modules() if getattr(mod, "use_batchnorm", True) is False for c in mod.children()
                  if isinstance(c, torch.nn.BatchNorm2d)}      # the fork's BN-free blocks keep (and never step) their BatchNorm modules
        bns = [mod for mod in model.modules() if isinstance(mod, torch.nn.BatchNorm2d) and mod not in unused]
        self.flat_nbt = torch.zeros(len(bns), device=dev, dtype=torch.int64)
        for i, bn in enumerate(bns):
            self.flat_nbt[i] = bn.num_batches_tracked
            bn.num_batches_tracked = self.flat_nbt[i]       # a 0-dim view: state_dict() keeps reporting it
            bn.__dict__["_erf_nbt_managed"] = True
        # tcgen05 operand images of all 3-tap convs, re-packed from the flat parameters by one launch per step
        self._pack_table = None
        if dev.type == "cuda" and not ops.EMULATED:
            blocks = [m for m in model.modules() if isinstance(m, _erfnet.non_bottleneck_1d) and m.chann in (16, 64, 128)]
            if blocks:
                self._pack_table = ops.Conv1dPackTable([c.weight.data for b in blocks for c in b._convs()])
                for i, b in enumerate(blocks):
                    b.__dict__["_prepacked"] = self._pack_table.pairs[4 * i:4 * i + 4]
        self._pending = list(self._bucket_params)
        self._launched = [False, False, False]
        self._works = []
        self._graph, self._step_dev = None, None
        self._coalesce = None       # N > 1 under graphs: gradient + loss all-reduce as one NCCL group (decided at first use)
        self._hooks_enabled = True
        self._seen = set()
        self.cuda = dev.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.cuda and self.world > 1) else None
        if self.world > 1:
            for p in model.parameters():
                p.register_post_accumulate_grad_hook(self._hook)
            # every rank starts from rank 0's weights (nn.DataParallel replicates GPU0's module each forward)
            dist.broadcast(self.flat_p, 0, group=self.pg)
            for b in model.buffers():
                dist.broadcast(b, 0, group=self.pg)

    # ---- hyper-parameters: the reference steps a LambdaLR every epoch (train/main.py:238-239) and the ImageNet schedule decays lr every 30
    # epochs (imagenet/main.py:276-286).  A captured graph cannot see host-side floats, so once a graph exists lr / weight decay live in a
    # two-float device tensor that the Adam / SGD kernels read; assigning trainer.lr (or set_lr) updates it, also between replays.
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        self._sync_hyper()

    @property
    def wd(self):
        return self._wd

    @wd.setter
    def wd(self, value):
        self._wd = float(value)
        self._sync_hyper()

    def set_lr(self, lr):
        """scheduler.step() of the reference: takes effect at the next step, graph replay included."""
        self.lr = lr

    def _sync_hyper(self):
        if self._hyper_dev is not None:
            self._hyper_dev.copy_(torch.tensor([self._lr, self._wd], dtype=torch.float32), non_blocking=False)

    @property
    def param_groups(self):
        """Read-only view in torch.optim's vocabulary (what `for g in optimizer.param_groups: g['lr']` loops expect to read); assign
        through ``trainer.lr`` / ``set_lr``."""
        g = {"params": self._params, "lr": self._lr, "weight_decay": self._wd}
        if self.sgd_momentum is not None:
            g.update(momentum=self.sgd_momentum, dampening=0, nesterov=False)
        else:
            g.update(betas=tuple(self.betas), eps=self.eps, amsgrad=False)
        return [g]

    # ---- checkpoints: the layout torch.optim.Adam / SGD produce for optimizer.state_dict() (train/main.py:229, 783 save and resume
    # checkpoint['optimizer']), so a reference checkpoint resumes here and vice versa -------------------------------------------------
    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for i, p in enumerate(self._params):
                o, n = self._offsets[p], p.numel()
                if self.sgd_momentum is not None:
                    state[i] = {"momentum_buffer": self.flat_m[o:o + n].view_as(p).clone()}
                else:
                    state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.flat_m[o:o + n].view_as(p).clone(),
                                "exp_avg_sq": self.flat_v[o:o + n].view_as(p).clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self._params)))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self._params):
            raise ValueError(f"FlatAdamTrainer.load_state_dict: {len(ids)} parameters in the checkpoint, {len(self._params)} in the model")
        g0 = groups[0]
        if "betas" in g0 and self.sgd_momentum is None:
            self.betas, self.eps = tuple(g0["betas"]), g0["eps"]
        self._lr, self._wd = float(g0["lr"]), float(g0.get("weight_decay", self._wd))
        state = sd.get("state", {})
        step = 0
        with torch.no_grad():
            self.flat_m.zero_()
            self.flat_v.zero_()
            for pos, i in enumerate(ids):
                st = state.get(i, state.get(str(i)))
                if not st:
                    continue
                p = self._params[pos]
                o, n = self._offsets[p], p.numel()
                if self.sgd_momentum is not None:
                    if st.get("momentum_buffer") is not None:
                        self.flat_m[o:o + n].copy_(st["momentum_buffer"].reshape(-1))
                        step = max(step, 1)
                else:
                    self.flat_m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self.flat_v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, int(float(st["step"])))
        self.step_count = int(sd.get("step_count", step))
        if self._step_dev is not None:
            self._step_dev.fill_(self.step_count)
        self._sync_hyper()

    # ---- gradient buckets ---------------------------------------------------------------------------------
    def _launch_bucket(self, b):
        if self._launched[b] or self.world == 1:
            return
        self._launched[b] = True
        lo, hi = self.bucket_slices[b]
        if lo is None:
            return
        buf = self.flat_g[lo:hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            if ops.SIDE is not None and ops.SIDE.used:
                self.comm_stream.wait_stream(ops.SIDE.stream)      # this bucket's weight gradients were accumulated on the side stream
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def ready(self, p):
        """Called by the blocks' backward once the kernels have accumulated p's gradient into p.grad (gradient sink)."""
        self._hook(p)

    def _hook(self, p):
        # reached both from autograd's post-accumulate hook and from the gradient sink: count each parameter once per step
        if not self._hooks_enabled or p in self._seen:
            return
        self._seen.add(p)
        b = self._param_bucket[p]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch_bucket(b)

    def _finish_buckets(self):
        for b in range(3):  # buckets holding parameters that received no gradient (encoder.output_conv) never count down
            self._launch_bucket(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = list(self._bucket_params)
        self._launched = [False, False, False]
        self._seen = set()

    # ---- one step ---------------------------------------------------------------------------------------------
    def _generic_loss_and_backward(self, images, targets):
        """Any net + any criterion: public forward, autograd through the criterion, kernels accumulate into the flat gradient views."""
        try:
            out = self.model(images)
        finally:
            _erfnet._PREPACKED_FRESH = False
        if self.world > 1:      # this package's criteria then normalise by the GLOBAL batch (see ops.LOSS_REDUCE), like the reference under DataParallel
            ops.LOSS_REDUCE = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            ops.LOSS_REDUCED = False
        try:
            loss = self.loss_fn(out, *targets)
        finally:
            ops.LOSS_REDUCE = None
        # a foreign criterion (plain torch ops) yields the rank's own mean: average those instead
        self._grad_scale = 1.0 if (self.world == 1 or ops.LOSS_REDUCED) else 1.0 / self.world
        self._backward(loss, None)
        self._finish_buckets()
        loss = loss.detach()
        if self.world > 1 and not ops.LOSS_REDUCED:
            loss = loss.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.pg)
            loss /= self.world
        return loss

    def loss_and_backward(self, images, labels, drop_scales=None):
        """images NCHW fp32, labels int64 [B,H,W] on the model's device.  Returns the GLOBAL-batch loss (0-dim tensor).
        With a ``loss_fn``: labels is the target tensor (or a tuple of them) that follows the net output in the criterion call."""
        self._begin_step()
        try:
            return self._loss_and_backward(images, labels, drop_scales)
        finally:
            self._end_step()

    def _begin_step(self):
        """Zero the flat gradient buffer and the step's scratch arena (BN batch sums of forward and backward, grid-barrier counters: ONE
        memset), step the BatchNorm counters, re-pack the tensor-core weight images."""
        # The weight re-pack comes FIRST and the memsets after it: a memset is a full serialisation point of the stream, so every kernel
        # launched later -- including the part of a programmatically dependent kernel that runs before its pdl_wait() -- starts after the
        # packed images are complete.  That is what lets the training convs fetch their weights before pdl_wait() (static_w).
        if self._pack_table is not None:
            if not self._pack_table.valid():
                raise RuntimeError("FlatAdamTrainer: a conv weight was re-allocated after the trainer took ownership of the parameters")
            self._pack_table.run()
            _erfnet._PREPACKED_FRESH = True
        if self.cuda and not ops.EMULATED and ops.FORK_JOIN and torch.cuda.is_current_stream_capturing():
            # the im2col convs' weight images (forward and data gradient) for this step, packed on a second stream next to the first layers
            convs = []
            for mod in self.model.modules():
                tc = mod.__dict__.get("_tc")
                if isinstance(tc, _erfnet._TcConv) and getattr(mod, "conv", None) is not None:
                    convs.append((tc, mod.conv.weight))
                for v in mod.__dict__.values():
                    if isinstance(v, _erfnet._PlainConv):
                        convs.append((v._tc, v._conv.weight))
            if convs:
                prec = self.model._precision() if hasattr(self.model, "_precision") else "bf16"
                _erfnet.prepack_begin(convs, torch.bfloat16 if prec == "bf16" else torch.float32)
        self.flat_g.zero_()
        if self.flat_nbt.numel():
            self.flat_nbt.add_(1)
        self.model.train()
        if self.cuda and not ops.EMULATED:
            if getattr(self, "_arena", None) is None:
                self._arena = ops.ScratchArena(self.flat_p.device)
            self._arena.reset()
            ops.ARENA = self._arena

    def _end_step(self):
        _erfnet._PREPACKED_FRESH = False
        _erfnet.prepack_end()
        ops.ARENA = None

    def _loss_and_backward(self, images, labels, drop_scales=None):
        if self.loss_fn is not None:
            return self._generic_loss_and_backward(images, labels if isinstance(labels, (tuple, list)) else (labels,))
        try:
            logits = self.model.forward_nhwc(images, drop_scales=drop_scales)
        finally:
            _erfnet._PREPACKED_FRESH = False
        c = self.num_classes
        w = self.criterion.weight
        labels = labels.contiguous()
        ops.check_ce_shapes(logits, labels, w, c, False)
        # pass 1: valid-pixel weight sum, all-reduced so that every rank normalises by the global denominator
        acc = torch.zeros(2, device=logits.device, dtype=torch.float32)
        wsum, lsum = acc[0:1], acc[1:2]
        ops.ce_weight_sum(labels, w, c, self.criterion.ignore_index, wsum)
        if self.world > 1:
            dist.all_reduce(wsum, op=dist.ReduceOp.SUM, group=self.pg)
        dlog = self._ce(logits, labels, wsum, lsum)
        self._backward(logits, dlog)
        self._finish_buckets()
        if self.world > 1:
            dist.all_reduce(lsum, op=dist.ReduceOp.SUM, group=self.pg)
        return (lsum / wsum).reshape(())

    def _ce(self, logits, labels, wsum, lsum):
        """Fused CE forward + dlogits.  The per-class sums of dlogits ARE the bias gradient of decoder.output_conv (the logits are its raw
        output): where the kernel supports it they are accumulated into that gradient here and the head's backward skips its own pass over
        the 151 MB of dlogits."""
        head = getattr(getattr(self.model, "decoder", None), "output_conv", None)
        colsum = None
        if (head is not None and getattr(head, "bias", None) is not None and head.bias.grad is not None and head.bias.numel() == self.num_classes
                and ops.ce_colsum_supported(logits)):
            colsum = head.bias.grad
            head.__dict__["_erf_bias_grad_done"] = True      # consumed by _PlainConv._bwd of this step
        return ops.ce_fwd_bwd(logits.detach(), labels, self.criterion.weight, self.num_classes, self.criterion.ignore_index, wsum, lsum, colsum=colsum)

    def _backward(self, logits, dlog):
        _erfnet._GRAD_SINK = self
        side = ops.WGRAD_SIDE_STREAM
        if side == "auto":
            side = self.cuda and torch.cuda.is_current_stream_capturing()
        if self.cuda and not ops.EMULATED and side:
            if getattr(self, "_side", None) is None:
                self._side = ops.SideQueue(self.flat_p.device)
            ops.SIDE = self._side
            if ops.SIDE_CAP[0] or ops.SIDE_CAP[1]:
                ops.set_launch_cap(*ops.SIDE_CAP)
        try:
            if dlog is None:
                logits.backward()          # generic path: `logits` is the scalar loss
            else:
                logits.backward(dlog)
        finally:
            _erfnet._GRAD_SINK = None
            if ops.SIDE is not None:
                ops.SIDE.join()            # the optimiser (and the gradient all-reduce) wait for the weight gradients
                ops.SIDE = None
                ops.set_launch_cap(0, 0)

    def optimizer_step(self):
        self.step_count += 1
        gs = getattr(self, "_grad_scale", 1.0) if self.loss_fn is not None else 1.0   # 1 when the loss was normalised by the global batch
        if self.sgd_momentum is not None:
            # the momentum buffer starts at zero, so momentum * buf + d on the first step equals torch's "buf = d" initialisation
            if self._hyper_dev is not None:
                ops.sgd_step_graph(self.flat_p, self.flat_g, self.flat_m, self._hyper_dev, self.sgd_momentum, gs)
            else:
                ops.sgd_step(self.flat_p, self.flat_g, self.flat_m, self.lr, self.sgd_momentum, self.wd, False, gs)
            _erfnet.bump_weights_epoch()
            return
        if self._step_dev is not None:  # CUDA-graph mode: step counter, learning rate and weight decay live on the device
            ops.adam_step_graph(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self._hyper_dev, self.betas[0], self.betas[1], self.eps,
                                self._step_dev, gs)
        else:
            ops.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                          self.step_count, gs)
        _erfnet.bump_weights_epoch()  # flat_p aliases the parameters; invalidate eval-mode packed weights

    def train_step(self, images, labels, drop_scales=None):
        if self._graph is not None and drop_scales is None:
            self._g_images.copy_(images, non_blocking=True)
            self._g_labels.copy_(labels, non_blocking=True)
            if isinstance(self._graph, tuple):
                # N > 1: two graph segments with the NCCL collectives launched eagerly around them.  The global valid-pixel weight sum
                # depends on the labels only, so it is reduced BEFORE the step and forward + loss + backward stay one graph.
                g12, g3 = self._graph
                self._g_acc.zero_()
                ops.ce_weight_sum(self._g_labels, self.criterion.weight, self.num_classes, self.criterion.ignore_index, self._g_acc[0:1])
                dist.all_reduce(self._g_acc[0:1], op=dist.ReduceOp.SUM, group=self.pg)
                g12.replay()                                                   # zero grads, forward, CE fwd+bwd, backward
                self._allreduce_grads_and_loss()
                g3.replay()                                                    # Adam, loss = lsum / wsum
            else:
                self._graph.replay()
            self.step_count += 1
            _erfnet.bump_weights_epoch()
            return self._g_loss
        loss = self.loss_and_backward(images, labels, drop_scales)
        self.optimizer_step()
        return loss

    def _allreduce_grads_and_loss(self):
        """Gradient buffer and loss sum between the two graph segments: one NCCL group (one launch) where the backend coalesces."""
        if self._coalesce is None:
            self._coalesce = dist.get_backend(self.pg) == "nccl" and hasattr(dist, "_coalescing_manager")
        if self._coalesce:
            try:
                with dist._coalescing_manager(group=self.pg, device=self.flat_p.device):
                    dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
                    dist.all_reduce(self._g_acc[1:2], op=dist.ReduceOp.SUM, group=self.pg)
                return
            except Exception:      # private API of torch.distributed: fall back to two launches for good
                self._coalesce = False
        dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
        dist.all_reduce(self._g_acc[1:2], op=dist.ReduceOp.SUM, group=self.pg)

    # ---- host-fed training: H2D of batch i+1 overlaps the step on batch i ------------------------------------------------
    def _stage(self, batch, k):
        """Enqueue the host->device copy of (images, labels) (pinned CPU tensors) into staging slot k on the copy stream."""
        images, labels = batch
        dev = self.flat_p.device
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(device=dev)
            self._stage_buf = [None, None]
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_used = [False, False]
        buf = self._stage_buf[k]
        if buf is None or buf[0].shape != images.shape or buf[1].shape != labels.shape or buf[1].dtype != labels.dtype or buf[0].dtype != images.dtype:
            # labels keep their dtype (int64 class maps, float regression targets of the fork nets); images may be fp32 NCHW or uint8 HWC
            buf = self._stage_buf[k] = (torch.empty(images.shape, device=dev, dtype=images.dtype), torch.empty(labels.shape, device=dev, dtype=labels.dtype))
            self._stage_used[k] = False
        with torch.cuda.stream(self._h2d_stream):
            if self._stage_used[k]:
                self._h2d_stream.wait_event(self._stage_free[k])      # the step that read this slot has been enqueued and finished
            buf[0].copy_(images, non_blocking=True)
            buf[1].copy_(labels, non_blocking=True)
            self._stage_ready[k].record(self._h2d_stream)

    def _step_from_stage(self, k):
        cur = torch.cuda.current_stream(self.flat_p.device)
        cur.wait_event(self._stage_ready[k])
        loss = self.train_step(*self._stage_buf[k])
        self._stage_free[k].record(cur)
        self._stage_used[k] = True
        return loss

    def train_stream(self, batches, host_losses=False):
        """Generator over an iterable of HOST batches (images fp32 NCHW, labels int64; pinned memory for asynchronous copies): yields
        the loss of each step, in order, one per batch.  The copy of the next batch runs on a side stream while the current step computes.

        host_losses=False: yields the step's loss TENSOR right after the step was enqueued; a loop that calls float() on it every step
        (train/main.py:392) waits for the step and leaves the GPU idle while the next one is being launched.
        host_losses=True: yields Python floats.  Every step's loss is copied to pinned host memory by an asynchronous device->host copy behind
        the step, and the value of step i is handed out after step i+1 has been ENQUEUED (the last one after the loop): the same values in the
        same order, every one read on the host, without that bubble."""
        it = iter(batches)
        nxt = next(it, None)
        k = 0
        if nxt is not None:
            self._stage(nxt, k)
        pending = None          # (pinned host slot, event) of the previous step's loss
        if host_losses and getattr(self, "_loss_host", None) is None:
            self._loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
        while nxt is not None:
            nxt2 = next(it, None)
            if nxt2 is not None:
                self._stage(nxt2, k ^ 1)
            loss = self._step_from_stage(k)
            if host_losses:
                slot, ev = self._loss_host[k], self._loss_evt[k]
                slot.copy_(loss.detach().reshape(1).float(), non_blocking=True)
                ev.record(torch.cuda.current_stream(self.flat_p.device))
                if pending is not None:
                    pending[1].synchronize()
                    yield float(pending[0][0])
                pending = (slot, ev)
            else:
                yield loss
            nxt, k = nxt2, k ^ 1
        if pending is not None:
            pending[1].synchronize()
            yield float(pending[0][0])

    # ---- CUDA graph of the whole step (forward, loss, backward, all-reduce, Adam): ~700 launches -> one replay ---------
    def capture_cuda_graph(self, images_shape, labels_shape, warmup=2, labels_dtype=torch.int64):
        """Capture one full training step for fixed shapes.  State touched by the warm-up steps is restored, so training
        semantics are unchanged; afterwards train_step() copies the batch into static buffers and replays the graph.
        labels_dtype: torch.float32 for the fork's regression targets."""
        dev = self.flat_p.device
        if self.loss_fn is not None and self.world > 1:
            raise RuntimeError("FlatAdamTrainer: graph capture of the generic (loss_fn) step is built for one GPU; run it eagerly at N > 1")
        self._g_images = torch.zeros(images_shape, device=dev, dtype=torch.float32)
        self._g_labels = torch.zeros(labels_shape, device=dev, dtype=labels_dtype)
        self._step_dev = torch.full((1,), self.step_count, device=dev, dtype=torch.int32)
        self._hyper_dev = torch.tensor([self._lr, self._wd], device=dev, dtype=torch.float32)
        snap = [t.clone() for t in (self.flat_p, self.flat_m, self.flat_v, self._step_dev)]
        bufs = [b.clone() for b in self.model.buffers()]
        count = self.step_count
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.loss_and_backward(self._g_images, self._g_labels)
                self.optimizer_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        # Parameters that receive their gradient through autograd (class_power, the folded 32-channel weights) own AccumulateGrad nodes that
        # remember the stream they were created on; a dead-but-uncollected graph from the steps above would hand those stale nodes to the
        # captured forward and make the engine synchronise an uncaptured stream with the capturing one.  Collect them now.
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        if self.world > 1:
            # NCCL collectives stay outside the graphs (8.27 MB of gradients + two scalars per step): capture the three
            # compute segments between them, sharing one memory pool so that activations saved by segment 1 stay alive
            self._hooks_enabled = False
            c, w = self.num_classes, self.criterion.weight
            self._g_acc = torch.zeros(2, device=dev, dtype=torch.float32)
            g12, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            ops.ce_weight_sum(self._g_labels, w, c, self.criterion.ignore_index, self._g_acc[0:1])   # any non-zero denominator for the capture
            with torch.cuda.graph(g12):
                self._begin_step()
                try:
                    try:
                        logits = self.model.forward_nhwc(self._g_images)
                    finally:
                        _erfnet._PREPACKED_FRESH = False
                    dlog = self._ce(logits, self._g_labels, self._g_acc[0:1], self._g_acc[1:2])
                    self._backward(logits, dlog)
                finally:
                    self._end_step()
            with torch.cuda.graph(g3, pool=g12.pool()):
                self.optimizer_step()
                loss = (self._g_acc[1:2] / self._g_acc[0:1]).reshape(())
            self._hooks_enabled = True
            graph = (g12, g3)
        else:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = self.loss_and_backward(self._g_images, self._g_labels)
                self.optimizer_step()
        for t, s_ in zip((self.flat_p, self.flat_m, self.flat_v, self._step_dev), snap):
            t.copy_(s_)
        for b, s_ in zip(self.model.buffers(), bufs):
            b.copy_(s_)
        self.step_count = count
        self._g_loss, self._graph = loss, graph
        return graph


def allreduce_confusion(ev, process_group=None):
    """Sum the per-rank iouEval confusion matrices (config 5: eval shards by batch; only 400 counters cross GPUs)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        cm = ev.confusion_matrix()
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=process_group)
        ev._cm = cm
        ev._dirty = True
    return ev
