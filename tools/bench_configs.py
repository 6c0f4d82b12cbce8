"""Throughput of the other BASELINE configs (parity-test cases, not bench lines): config 3 (ImageNet encoder pre-training step, 224 x 224) and
config 4 (fork ladder regression net, 512 x 1024, masked L1), one GPU, eager launches, CUDA events.  For DESIGN.md section 7."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import erfnet_imagenet as EI  # noqa: E402
from erfnet_pytorch_b200 import evaluation_functions as EV  # noqa: E402
from erfnet_pytorch_b200 import net_regression_ladder as NL  # noqa: E402
from erfnet_pytorch_b200.optim import FlatSGD  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


res = {}
# ---- config 3: erfnet_imagenet.ERFNet(1000), CE, SGD(0.1, 0.9, 1e-4)
B = int(os.environ.get("IMAGENET_BATCH", "64"))
m = EI.ERFNet(1000)
m.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=51))
m = m.cuda().train()
opt = FlatSGD(m.parameters(), 0.1, momentum=0.9, weight_decay=1e-4)
crit = EI.CrossEntropyLoss()
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")


def step3():
    loss = crit(m(x), y)
    opt.zero_grad()
    loss.backward()
    opt.step()


ms = timed(step3)
res["config3_imagenet_train"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "note": "224x224: widths 56 / 28 in the encoder: ragged 128-pixel units on the tcgen05 kernels since round 2 (CUDA-core kernels in round 1)"}
del opt
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402
tr3 = FlatAdamTrainer(m, lr=0.1, weight_decay=1e-4, loss_fn=lambda out, t: crit(out, t), sgd_momentum=0.9)
ms = timed(lambda: tr3.train_step(x, y))
res["config3_imagenet_train_flat"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "note": "FlatAdamTrainer(loss_fn=CE, sgd_momentum=0.9), eager"}
tr3.capture_cuda_graph(tuple(x.shape), tuple(y.shape))
ms = timed(lambda: tr3.train_step(x, y))
res["config3_imagenet_train_flat_graph"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "note": "the same step replayed from one CUDA graph"}
del tr3
m.eval()
with torch.no_grad():
    ms = timed(lambda: m(x))
res["config3_imagenet_eval"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3}
del m, x, y
torch.cuda.empty_cache()

# ---- config 4: net_regression_ladder.Net(), L1LossMasked, Adam as train/main.py:43-46
B = 6
net = NL.Net()
net.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=31))
net = net.cuda().train()
opt = torch.optim.Adam(net.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
crit = EV.L1LossMasked()
x = torch.rand(B, 3, 512, 1024, device="cuda")
tgt = torch.rand(B, 512, 1024, device="cuda") * 2
tgt[torch.rand(B, 512, 1024, device="cuda") < 0.7] = -1.0


def step4():
    loss = crit(net(x), tgt)
    opt.zero_grad()
    loss.backward()
    opt.step()


ms = timed(step4)
res["config4_ladder_train"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "tc_frac": B / ms * 1e3 * 225.2 / 1e3 / 1427.3,
                               "note": "eager launches through autograd + torch.optim.Adam (no flat trainer / CUDA graph for the fork nets)"}
del opt
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402
tr = FlatAdamTrainer(net, loss_fn=lambda out, t: crit(out, t))
ms = timed(lambda: tr.train_step(x, tgt))
res["config4_ladder_train_flat"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "note": "FlatAdamTrainer(loss_fn=...): gradient sink + one Adam kernel, eager"}
tr.capture_cuda_graph(tuple(x.shape), tuple(tgt.shape), labels_dtype=torch.float32)
ms = timed(lambda: tr.train_step(x, tgt))
res["config4_ladder_train_flat_graph"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "tc_frac": B / ms * 1e3 * 225.2 / 1e3 / 1427.3,
                                          "note": "the same step replayed from one CUDA graph"}
net.eval()
with torch.no_grad():
    ms = timed(lambda: net(x))
res["config4_ladder_eval"] = {"batch": B, "ms_per_step": ms, "images_per_s": B / ms * 1e3, "tc_frac": B / ms * 1e3 * 75.05 / 1e3 / 1427.3}
# ---- the fork's own frame size: 640 x 480: widths 320 / 160 / 80 are not multiples of 128 -- ragged last units since round 2
del tr
net.train()
B2 = 6
x2 = torch.rand(B2, 3, 480, 640, device="cuda")
tgt2 = torch.rand(B2, 480, 640, device="cuda") * 2
tgt2[torch.rand(B2, 480, 640, device="cuda") < 0.7] = -1.0
tr2 = FlatAdamTrainer(net, loss_fn=lambda out, t: crit(out, t))
tr2.capture_cuda_graph(tuple(x2.shape), tuple(tgt2.shape), labels_dtype=torch.float32)
ms = timed(lambda: tr2.train_step(x2, tgt2))
res["config4_ladder_train_480x640_graph"] = {"batch": B2, "ms_per_step": ms, "images_per_s": B2 / ms * 1e3, "note": "the reference fork's frame size (train/main.py crops to 640x480)"}
net.eval()
with torch.no_grad():
    ms = timed(lambda: net(x2))
res["config4_ladder_eval_480x640"] = {"batch": B2, "ms_per_step": ms, "images_per_s": B2 / ms * 1e3}
# ---- I/O side: uint8 HWC ingest (ToTensor + layout) at the bench shape, 3 B in + 8 B out per pixel
from erfnet_pytorch_b200 import ops  # noqa: E402
img = torch.randint(0, 256, (24, 512, 1024, 3), device="cuda", dtype=torch.uint8)
ms = timed(lambda: ops.ingest_u8(img, 4, torch.bfloat16), steps=20)
res["ingest_u8_bf16"] = {"batch": 24, "ms_per_step": ms, "images_per_s": 24 / ms * 1e3, "gbytes_per_s": img.numel() / 3 * 11 / ms / 1e6,
                         "note": "algorithmic 11 B/pixel (3 in + 8 out); measured HBM peak 6445 GB/s"}
print(json.dumps(res))
