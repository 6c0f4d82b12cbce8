"""Flat-buffer optimiser pieces outside the Cityscapes Adam trainer (parallel.FlatAdamTrainer):

  * ``FlatSGD``      -- ``torch.optim.SGD(params, lr, momentum, weight_decay)`` of imagenet/main.py:83-85 as ONE kernel over one flat fp32
                        parameter / gradient / momentum buffer (the parameters become views of the flat buffer);
  * ``MeanTeacher``  -- the teacher update of train/main.py:130-142, 384-390 (``copyWeightsToModelWithDiscount``) as ONE kernel over the
                        student's and the teacher's flat buffers.
Everything arithmetic is a kernel of liberfnet_b200.so (csrc/tail_io.cu); torch only owns the memory."""
from __future__ import annotations

import torch

from . import erfnet as _E
from . import ops


def flatten_parameters(params, like=None):
    """Move ``params`` into one flat fp32 buffer (each parameter becomes a view of it) and return the buffer.
    If they already are consecutive-or-not views of one storage, that storage's covering slice is returned unchanged.
    ``like``: a list of parameters of a model that is already flat -- the new buffer then uses the SAME offsets, element for element."""
    params = list(params)
    if not params:
        raise ValueError("no parameters")
    base = params[0].data.untyped_storage().data_ptr()
    if like is None and all(p.data.untyped_storage().data_ptr() == base and p.data.is_contiguous() for p in params) and len(params) > 1:
        lo = min(p.data.storage_offset() for p in params)
        hi = max(p.data.storage_offset() + p.numel() for p in params)
        if hi - lo == sum(p.numel() for p in params):
            whole = torch.empty(0, device=params[0].device, dtype=torch.float32).set_(params[0].data.untyped_storage(), lo, (hi - lo,), (1,))
            return whole
    if like is not None:
        like = list(like)
        lo = min(q.data.storage_offset() for q in like)
        offs = [q.data.storage_offset() - lo for q in like]
        total = max(o + q.numel() for o, q in zip(offs, like))
    else:
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += p.numel()
    flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
    for p, o in zip(params, offs):
        flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
        p.data = flat[o:o + p.numel()].view_as(p)
    return flat


class FlatSGD:
    """Drop-in for ``torch.optim.SGD(model.parameters(), lr, momentum=..., weight_decay=...)`` (no dampening, no Nesterov -- what
    imagenet/main.py:83-85 uses): ``zero_grad()``, ``step()``, ``param_groups`` (lr / weight_decay are read at every step, so
    ``adjust_learning_rate`` works), ``state_dict()`` / ``load_state_dict()`` for checkpoints (l.89-97, 146-152)."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        self.flat_p = flatten_parameters(self.params)
        self.flat_g = torch.zeros_like(self.flat_p)
        lo = min(p.data.storage_offset() for p in self.params)
        for p in self.params:
            o = p.data.storage_offset() - lo
            p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        self.flat_buf = torch.zeros_like(self.flat_p) if momentum != 0 else None
        self.param_groups = [{"params": self.params, "lr": lr, "momentum": momentum, "weight_decay": weight_decay}]
        self.steps = 0

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()          # the .grad views stay in place: autograd / the gradient sink accumulate into them

    def step(self):
        g = self.param_groups[0]
        ops.sgd_step(self.flat_p, self.flat_g, self.flat_buf, float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]), self.steps == 0)
        self.steps += 1
        _E.bump_weights_epoch()      # the parameters changed in place: invalidate eval-mode packed weights

    def state_dict(self):
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        return {"param_groups": [g], "steps": self.steps, "momentum_buffer": None if self.flat_buf is None else self.flat_buf.clone()}

    def load_state_dict(self, sd):
        self.param_groups[0].update(sd["param_groups"][0])
        self.steps = sd["steps"]
        if self.flat_buf is not None and sd.get("momentum_buffer") is not None:
            self.flat_buf.copy_(sd["momentum_buffer"])


class MeanTeacher:
    """train/main.py:130-142: ``copyWeightsToModelNoGrad`` (which, in the reference, only switches ``requires_grad`` off -- its
    ``target = source.clone()`` rebinds a local name and copies nothing) and ``copyWeightsToModelWithDiscount``:
    teacher = teacher * (1 - discount) + discount * student, over all parameters, here one kernel launch."""

    def __init__(self, model_student, model_teacher):
        sp, tp = list(model_student.parameters()), list(model_teacher.parameters())
        if len(sp) != len(tp) or any(a.shape != b.shape for a, b in zip(sp, tp)):
            raise RuntimeError("MeanTeacher: student and teacher must have the same parameters")
        self.flat_student = flatten_parameters(sp)
        self.flat_teacher = flatten_parameters(tp, like=sp)
        for target in tp:
            target.requires_grad = False

    def update(self, discount_factor):
        ops.ema_update(self.flat_teacher, self.flat_student, float(discount_factor))
        _E.bump_weights_epoch()
