"""Mean-teacher training iterations (train/main.py:297-392 with --mean-teacher) of our net_regression_small.Net pair
against the reference's own run (tests/golden/mean_teacher.npz): student / teacher / consistency losses per step, parameters afterwards."""
import os

import numpy as np
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from erfnet_pytorch_b200 import net_regression_small
from erfnet_pytorch_b200.optim import MeanTeacher
from oracle import erfnet_oracle as O
from oracle.make_golden import mean_teacher_inputs, mean_teacher_step
from tests.test_host_logic import assert_rel


def check_mean_teacher_training(golden_dir, device, precision="fp32", tol=2e-4):
    g = np.load(os.path.join(golden_dir, "mean_teacher.npz"))
    student, teacher = net_regression_small.Net(), net_regression_small.Net()
    shapes = {k: tuple(v.shape) for k, v in student.state_dict().items()}
    student.load_state_dict(O.make_state_dict(shapes, seed=41))
    teacher.load_state_dict(O.make_state_dict(shapes, seed=43))
    student.precision = teacher.precision = precision
    student, teacher = student.to(device).train(), teacher.to(device).train()
    mt = MeanTeacher(student, teacher)                     # flat buffers; teacher.requires_grad = False
    opt = torch.optim.Adam(list(student.parameters()), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
    x1, x2, tgt = (t.to(device) for t in mean_teacher_inputs())
    losses = [mean_teacher_step(student, teacher, EV, opt, x1, x2, tgt, 0.5, mt.update) for _ in range(3)]
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=tol)
    for k in ("decoder.scalar_output_conv.weight", "encoder.initial_block.conv.bias"):
        assert_rel(student.get_parameter(k).detach().cpu().numpy(), g["student_" + k], tol, "student " + k)
        assert_rel(teacher.get_parameter(k).detach().cpu().numpy(), g["teacher_" + k], tol, "teacher " + k)
