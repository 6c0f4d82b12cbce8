"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE (imported by path from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these files -- outputs of the
reference's own modules on deterministic inputs/weights -- are what pins oracle/erfnet_oracle.py.
Inputs and weights are NOT stored: they are regenerated from (seed, key) by oracle.erfnet_oracle.make_state_dict and
``golden_input`` below, so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import erfnet_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def golden_input(shape, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(shape, generator=g)


def golden_labels(shape, seed, n_classes=20, ignore=-1, p_ignore=0.05):
    g = torch.Generator().manual_seed(2000 + seed)
    lab = torch.randint(0, n_classes, shape, generator=g)
    drop = torch.rand(shape, generator=g) < p_ignore
    return torch.where(drop, torch.full_like(lab, ignore), lab)


def main():
    sys.path.insert(0, os.path.join(REF, "eval"))
    import erfnet as ref_erfnet  # /root/reference/eval/erfnet.py
    import iouEval as ref_iou  # /root/reference/eval/iouEval.py
    sys.path.insert(0, os.path.join(REF, "train"))
    import evaluation_functions as ref_loss  # /root/reference/train/evaluation_functions.py

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    shapes = O.erfnet_param_shapes(20)
    sd = O.make_state_dict(shapes, seed=0)

    # ---- state_dict layout (key names, shapes, dtypes) of the reference ERFNet(20)
    model = ref_erfnet.ERFNet(20)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), "key order/layout mismatch"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    model.load_state_dict(sd, strict=True)
    with open(os.path.join(OUT, "state_dict_keys.txt"), "w") as f:
        for k, v in ref_sd.items():
            f.write(f"{k} {tuple(v.shape)} {str(v.dtype).replace('torch.', '')}\n")

    # ---- whole net, eval mode
    model.eval()
    x = golden_input((1, 3, 64, 128), 0)
    with torch.no_grad():
        logits = model(x)
        enc = model(x, only_encode=True)
    np.savez_compressed(os.path.join(OUT, "erfnet_eval_64x128.npz"), logits=logits.numpy(), enc_logits=enc.numpy())

    # ---- blocks, eval mode (each (C, dilation) of the net) and train mode (BN batch stats, fixed dropout mask off)
    blocks = {}
    for name, (c, d, hw) in {"nb64_d1": (64, 1, (8, 24)), "nb128_d2": (128, 2, (8, 12)), "nb128_d16": (128, 16, (18, 20)),
                             "nb16_d1": (16, 1, (16, 24)), "nb128_d4": (128, 4, (8, 16)), "nb128_d8": (128, 8, (10, 18))}.items():
        m = ref_erfnet.non_bottleneck_1d(c, 0.0, d)
        bsh = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        bsd = O.make_state_dict({"blk." + k: s for k, s in bsh.items()}, seed=3)
        m.load_state_dict({k[4:]: v for k, v in bsd.items()}, strict=True)
        xb = golden_input((2, c) + hw, 10 + d) - 0.3
        m.eval()
        with torch.no_grad():
            blocks[name + "_eval"] = m(xb).numpy()
        m.train()
        xb.requires_grad_(True)
        yb = m(xb)
        gy = golden_input(tuple(yb.shape), 77) - 0.5
        yb.backward(gy)
        blocks[name + "_train"] = yb.detach().numpy()
        blocks[name + "_train_dx"] = xb.grad.numpy()
        blocks[name + "_train_dw_conv1x3_2"] = m.conv1x3_2.weight.grad.numpy()
        blocks[name + "_train_dw_conv3x1_1"] = m.conv3x1_1.weight.grad.numpy()
        blocks[name + "_train_dgamma1"] = m.bn1.weight.grad.numpy()
        blocks[name + "_train_rm2"] = m.bn2.running_mean.numpy().copy()
        blocks[name + "_train_rv2"] = m.bn2.running_var.numpy().copy()
    for name, (ci, co, hw) in {"down3_16": (3, 16, (16, 32)), "down16_64": (16, 64, (8, 16)), "down64_128": (64, 128, (8, 8))}.items():
        m = ref_erfnet.DownsamplerBlock(ci, co)
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=4)
        m.load_state_dict({k[4:]: v for k, v in bsd.items()}, strict=True)
        xb = golden_input((2, ci) + hw, 20 + ci) - 0.3
        m.eval()
        with torch.no_grad():
            blocks[name + "_eval"] = m(xb).numpy()
        m.train()
        xb.requires_grad_(True)
        yb = m(xb)
        gy = golden_input(tuple(yb.shape), 78) - 0.5
        yb.backward(gy)
        blocks[name + "_train"] = yb.detach().numpy()
        blocks[name + "_train_dx"] = xb.grad.numpy()
        blocks[name + "_train_dw"] = m.conv.weight.grad.numpy()
    for name, (ci, co, hw) in {"up128_64": (128, 64, (4, 8)), "up64_16": (64, 16, (8, 8))}.items():
        m = ref_erfnet.UpsamplerBlock(ci, co)
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=5)
        m.load_state_dict({k[4:]: v for k, v in bsd.items()}, strict=True)
        xb = golden_input((2, ci) + hw, 30 + ci) - 0.3
        m.eval()
        with torch.no_grad():
            blocks[name + "_eval"] = m(xb).numpy()
        m.train()
        xb.requires_grad_(True)
        yb = m(xb)
        gy = golden_input(tuple(yb.shape), 79) - 0.5
        yb.backward(gy)
        blocks[name + "_train"] = yb.detach().numpy()
        blocks[name + "_train_dx"] = xb.grad.numpy()
        blocks[name + "_train_dw"] = m.conv.weight.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **blocks)

    # ---- one full training step of ERFNet(20): train-mode forward (dropout disabled by p=0 so that no RNG is
    # involved), CrossEntropyLoss2d, backward, Adam (train/main.py:43-46,213)
    model = ref_erfnet.ERFNet(20)
    model.load_state_dict(sd, strict=True)
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    model.train()
    xt = golden_input((2, 3, 32, 64), 5)
    yt = golden_labels((2, 32, 64), 5)
    crit = ref_loss.CrossEntropyLoss2d()
    opt = torch.optim.Adam(model.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
    losses = []
    for _ in range(2):
        opt.zero_grad()
        out = model(xt)
        loss = crit(out, yt)
        loss.backward()
        if not losses:
            g_first = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
            logits_train = out.detach().clone()
        opt.step()
        losses.append(float(loss))
    keys = ["decoder.output_conv.weight", "decoder.layers.5.conv1x3_2.weight", "decoder.layers.3.conv.weight", "decoder.layers.0.bn.weight",
            "encoder.layers.14.conv3x1_2.weight", "encoder.layers.6.conv.weight", "encoder.layers.3.bn1.bias",
            "encoder.layers.1.conv1x3_1.bias", "encoder.initial_block.conv.weight"]
    tr = {"losses": np.array(losses, np.float64), "logits_step0": logits_train.numpy()}
    for k in keys:
        tr["grad0/" + k] = g_first[k].numpy()
        tr["param2/" + k] = dict(model.named_parameters())[k].detach().numpy()
    tr["no_grad_keys"] = np.array([k for k, p in model.named_parameters() if p.grad is None])
    tr["rm/encoder.layers.9.bn2.running_mean"] = model.state_dict()["encoder.layers.9.bn2.running_mean"].numpy()
    tr["rv/decoder.layers.3.bn.running_var"] = model.state_dict()["decoder.layers.3.bn.running_var"].numpy()
    np.savez_compressed(os.path.join(OUT, "train_step_32x64.npz"), **tr)

    # ---- loss KATs: unweighted (fork-exact) CrossEntropyLoss2d
    lg = (golden_input((2, 20, 16, 24), 8) - 0.5) * 6
    lb = golden_labels((2, 16, 24), 8)
    lg.requires_grad_(True)
    l = crit(lg, lb)
    l.backward()
    wv = torch.linspace(0.5, 1.5, 20)
    lw = torch.nn.functional.cross_entropy(lg.detach(), lb, weight=wv, ignore_index=-1)  # the extension's contract
    np.savez_compressed(os.path.join(OUT, "ce_loss.npz"), loss=np.float64(l.item()), dlogits=lg.grad.numpy(),
                        loss_weighted=np.float64(lw.item()))

    # ---- iouEval KATs (eval/iouEval.py): label-map input, one-hot input, nClasses<=ignoreIndex branch
    kat = {}
    g = torch.Generator().manual_seed(9)
    for name, (n, shape) in {"c20": (20, (3, 1, 33, 47)), "c20_big": (20, (2, 1, 128, 256)), "c5_noignore": (5, (2, 1, 16, 16))}.items():
        p = torch.randint(0, n, shape, generator=g)
        t = torch.randint(0, n, shape, generator=g)
        ev = ref_iou.iouEval(n)
        ev.addBatch(p, t)
        ev.addBatch(t.flip(0), t)  # a second batch accumulates
        mean, per = ev.getIoU()
        kat[name + "_tp"], kat[name + "_fp"], kat[name + "_fn"] = ev.tp.numpy(), ev.fp.numpy(), ev.fn.numpy()
        kat[name + "_mean"], kat[name + "_iou"] = np.float64(mean), per.numpy()
        kat[name + "_pred"], kat[name + "_gt"] = p.numpy().astype(np.uint8), t.numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "iou_kat.npz"), **kat)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
