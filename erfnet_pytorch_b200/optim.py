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


def _check_views(params, flat, who):
    """Every parameter must still be a view INTO `flat` (same storage, inside its range): Module.to() / a trainer that takes ownership of
    the parameters afterwards re-allocates them, and a kernel over the old flat buffer would silently update orphaned memory."""
    base = flat.untyped_storage().data_ptr()
    lo, hi = flat.storage_offset(), flat.storage_offset() + flat.numel()
    for p in params:
        d = p.data
        if d.untyped_storage().data_ptr() != base or not (lo <= d.storage_offset() and d.storage_offset() + d.numel() <= hi):
            raise RuntimeError(f"{who}: a parameter no longer aliases the flat buffer captured at construction (the model was moved, or another "
                               "trainer took ownership of its parameters afterwards) -- build this object last")


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
        self._check_alias()
        g = self.param_groups[0]
        ops.sgd_step(self.flat_p, self.flat_g, self.flat_buf, float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]), self.steps == 0)
        self.steps += 1
        _E.bump_weights_epoch()      # the parameters changed in place: invalidate eval-mode packed weights

    def _check_alias(self):
        _check_views(self.params, self.flat_p, "FlatSGD")

    def state_dict(self):
        """torch.optim.SGD's layout ({"state": {i: {"momentum_buffer"}}, "param_groups": [{..., "params": [0..n-1]}]}), which is what
        imagenet/main.py:89-97, 146-152 saves and resumes; "steps" is an extra key torch ignores."""
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g.update(dampening=0, nesterov=False, params=list(range(len(self.params))))
        state = {}
        if self.flat_buf is not None and self.steps > 0:
            lo = min(p.data.storage_offset() for p in self.params)
            for i, p in enumerate(self.params):
                o = p.data.storage_offset() - lo
                state[i] = {"momentum_buffer": self.flat_buf[o:o + p.numel()].view_as(p).clone()}
        return {"state": state, "param_groups": [g], "steps": self.steps}

    def load_state_dict(self, sd):
        g = {k: v for k, v in sd["param_groups"][0].items() if k in ("lr", "momentum", "weight_decay")}
        self.param_groups[0].update(g)
        if "momentum_buffer" in sd:      # this class's round-1 format: one flat buffer
            self.steps = sd["steps"]
            if self.flat_buf is not None and sd.get("momentum_buffer") is not None:
                self.flat_buf.copy_(sd["momentum_buffer"])
            return
        state = sd.get("state", {})
        ids = list(sd["param_groups"][0].get("params", range(len(self.params))))
        if len(ids) != len(self.params):
            raise ValueError(f"FlatSGD.load_state_dict: {len(ids)} parameters in the checkpoint, {len(self.params)} in the model")
        seen = False
        if self.flat_buf is not None:
            self.flat_buf.zero_()
            lo = min(p.data.storage_offset() for p in self.params)
            for pos, i in enumerate(ids):
                st = state.get(i, state.get(str(i)))
                if st and st.get("momentum_buffer") is not None:
                    p = self.params[pos]
                    o = p.data.storage_offset() - lo
                    self.flat_buf[o:o + p.numel()].copy_(st["momentum_buffer"].reshape(-1))
                    seen = True
        self.steps = int(sd.get("steps", 1 if seen else 0))


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
        self._student, self._teacher = model_student, model_teacher
        for target in tp:
            target.requires_grad = False

    def update(self, discount_factor):
        # Construction order matters: a FlatAdamTrainer built AFTER this object re-creates the student's parameters in its own
        # flat buffer, and Module.to() / .cuda() re-allocates storage; the buffers captured here would then be orphans averaged in silence.
        _check_views(self._student.parameters(), self.flat_student, "MeanTeacher (student)")     # the models' CURRENT parameters
        _check_views(self._teacher.parameters(), self.flat_teacher, "MeanTeacher (teacher)")
        ops.ema_update(self.flat_teacher, self.flat_student, float(discount_factor))
        _E.bump_weights_epoch()
