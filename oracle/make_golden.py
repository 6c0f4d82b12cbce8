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
    fork_losses(os.path.join(OUT, "fork_losses.npz"))
    blocks_nobn(os.path.join(OUT, "blocks_nobn.npz"))
    ladder_net(os.path.join(OUT, "ladder_net.npz"))
    fork_nets(os.path.join(OUT, "fork_nets.npz"))
    imagenet_net(os.path.join(OUT, "imagenet_net.npz"))
    mean_teacher(os.path.join(OUT, "mean_teacher.npz"))
    erfnet_nobn(os.path.join(OUT, "erfnet_nobn.npz"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


NOBN_NB = {"nb64_d1_drop": (64, 1, (8, 24), 0.3, True), "nb128_d4": (128, 4, (8, 16), 0.1, False), "nb16_d1": (16, 1, (16, 24), 0.0, True),
           "nb128_d16_drop": (128, 16, (18, 20), 0.1, True)}
NOBN_DOWN = {"down3_16": (3, 16, (16, 32)), "down16_64": (16, 64, (8, 16)), "down64_128": (64, 128, (8, 8))}
NOBN_UP = {"up128_64": (128, 64, (4, 8)), "up64_16": (64, 16, (8, 8))}


def blocks_nobn(path):
    """The fork's BN-free blocks (train/erfnet_blocks.py, use_batchnorm=False default) in train mode: output, dx, dW, db.
    Dropout masks are drawn under a fixed seed and stored ([N,C] multipliers) so the tests can inject them."""
    sys.path.insert(0, os.path.join(REF, "train"))
    import erfnet_blocks as ref_blocks
    res = {}

    def load(m, seed):
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed)
        m.load_state_dict({k[4:]: v for k, v in bsd.items()}, strict=True)
        m.train()

    for name, (c, d, hw, p, use_dropout) in NOBN_NB.items():
        m = ref_blocks.non_bottleneck_1d(c, p, d, use_dropout=use_dropout)
        load(m, 13)
        xb = (golden_input((2, c) + hw, 40 + d) - 0.3).requires_grad_(True)
        torch.manual_seed(123)
        yb = m(xb)
        if use_dropout and p != 0:
            torch.manual_seed(123)
            res[name + "_drop"] = torch.nn.functional.dropout2d(torch.ones(2, c, 1, 1), p, True).reshape(2, c).numpy()
        yb.backward(golden_input(tuple(yb.shape), 87) - 0.5)
        res[name + "_y"], res[name + "_dx"] = yb.detach().numpy(), xb.grad.numpy()
        for cn in ("conv3x1_1", "conv1x3_1", "conv3x1_2", "conv1x3_2"):
            res[f"{name}_db_{cn}"] = getattr(m, cn).bias.grad.numpy()
            if c <= 64 or cn in ("conv3x1_1", "conv1x3_2"):  # keeps the fixture small
                res[f"{name}_dw_{cn}"] = getattr(m, cn).weight.grad.numpy()
        assert m.bn1.weight.grad is None
    for table, cls, seed, gseed in ((NOBN_DOWN, ref_blocks.DownsamplerBlock, 14, 88), (NOBN_UP, ref_blocks.UpsamplerBlock, 15, 89)):
        for name, (ci, co, hw) in table.items():
            m = cls(ci, co)
            load(m, seed)
            xb = (golden_input((2, ci) + hw, 50 + ci) - 0.3).requires_grad_(True)
            yb = m(xb)
            yb.backward(golden_input(tuple(yb.shape), gseed) - 0.5)
            res[name + "_y"], res[name + "_dx"] = yb.detach().numpy(), xb.grad.numpy()
            res[name + "_dw"], res[name + "_db"] = m.conv.weight.grad.numpy(), m.conv.bias.grad.numpy()
    # DecoderBlock / DecoderBlockSmall (train/erfnet_blocks.py:114-135): upsampler + 1-2 nb1d blocks chained
    for name, cls in (("decblock64_16", ref_blocks.DecoderBlock), ("decblocksmall64_16", ref_blocks.DecoderBlockSmall)):
        m = cls(64, 16)
        load(m, 16)
        xb = (golden_input((2, 64, 6, 8), 61) - 0.3).requires_grad_(True)
        yb = m(xb)
        yb.backward(golden_input(tuple(yb.shape), 90) - 0.5)
        res[name + "_y"], res[name + "_dx"] = yb.detach().numpy(), xb.grad.numpy()
        res[name + "_dw_up"] = m.layers[0].conv.weight.grad.numpy()
        res[name + "_db_last"] = m.layers[-1].conv1x3_2.bias.grad.numpy()
    np.savez_compressed(path, **res)


def ladder_net(path):
    """train/net_regression_ladder.Net (BASELINE config 4): state_dict layout, eval outputs of the 1- and 2-channel heads and one
    training step's loss + a few gradients (L1LossMasked on a 70 %-unlabelled target) at 64 x 128."""
    import contextlib
    import io
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(REF, "train"))
    with contextlib.redirect_stdout(io.StringIO()):
        import evaluation_functions as ref_ev
        import net_regression_ladder as ref_net
        nets = {"reg": ref_net.Net(), "lik": ref_net.Net(likelihood_loss=True)}
    res = {}
    x = golden_input((2, 3, 64, 128), 71)
    tgt = golden_input((2, 64, 128), 72) * 2.0
    tgt[golden_input((2, 64, 128), 73) < 0.7] = -1.0
    for name, m in nets.items():
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = O.make_state_dict(shapes, seed=31)
        m.load_state_dict(sd, strict=True)
        if name == "reg":
            with open(os.path.join(OUT, "ladder_net_keys.txt"), "w") as f:
                for k, sh in shapes.items():
                    f.write(f"{k} {sh}\n")
        m.train()   # BN is off and dropout is off for softmax_classes = 0: train == eval arithmetic
        y = m(x)
        res[name + "_out"] = y.detach().numpy()
        if name == "reg":
            loss = ref_ev.L1LossMasked()(y, tgt)
        else:
            loss = ref_ev.LogLikelihoodLossMasked(opt_eps=0.05)(y[:, 0], y[:, 1].pow(2), tgt)   # train/main.py:356-365; eps 0.05 (not 1e-4) keeps the KAT well conditioned
        loss.backward()
        res[name + "_loss"] = np.float64(loss.item())
        for k in ("encoder.initial_block.conv.weight", "encoder.block2.5.conv3x1_2.weight", "decoder.ladder_block_1.conv.weight",
                  "decoder.ladder_block_2.layer.conv1x3_1.bias", "decoder.scalar_output_conv.weight", "decoder.scalar_decoder_2.layers.0.conv.bias"):
            mod, attr = k.rsplit(".", 1)
            res[f"{name}_g_{k}"] = getattr(m.get_submodule(mod), attr).grad.numpy()
    # class-probability variant (softmax_classes = 6): SoftMaxConv head + class_power parameters + L1LossClassProbMasked / the likelihood
    # version.  Dropout is forced off (p = 0 everywhere) so that train mode is deterministic.
    with contextlib.redirect_stdout(io.StringIO()):
        cls = {"cls": ref_net.Net(softmax_classes=6, spread_class_power=True, late_dropout_prob=0.0),
               "clslik": ref_net.Net(softmax_classes=6, likelihood_loss=True, spread_class_power=True, late_dropout_prob=0.0)}
    for name, m in cls.items():
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = O.make_state_dict(shapes, seed=33)
        sd["class_power"] = torch.linspace(0.0, 2.0, 6).view(1, 6, 1, 1)
        if "class_power_var" in sd:
            sd["class_power_var"] = torch.linspace(0.3, 0.8, 6).view(1, 6, 1, 1)
        m.load_state_dict(sd, strict=True)
        m.train()
        out = m(x)
        if name == "cls":
            prob, power = out
            loss = ref_ev.L1LossClassProbMasked()(prob, power, tgt)
        else:
            prob, power, var = out
            loss = ref_ev.LogLikelihoodLossClassProbMasked(opt_eps=0.05)(prob, power, var, tgt)
        loss.backward()
        res[name + "_prob"], res[name + "_loss"] = prob.detach().numpy(), np.float64(loss.item())
        for k in ("class_power", "decoder.scalar_output_conv.convolution.weight", "decoder.ladder_block_1.conv.weight", "encoder.block1.3.conv1x3_1.weight") + (("class_power_var",) if name == "clslik" else ()):
            if "." in k:
                mod, attr = k.rsplit(".", 1)
                gten = getattr(m.get_submodule(mod), attr).grad
            else:
                gten = getattr(m, k).grad
            res[f"{name}_g_{k}"] = gten.numpy()
    np.savez_compressed(path, **res)


FORK_NETS = {   # name -> (reference module, Net kwargs, head kind)
    "regression": ("net_regression", dict(), "reg"),
    "regression_lik": ("net_regression", dict(likelihood_loss=True), "lik"),
    "regression_small": ("net_regression_small", dict(), "reg"),
    "regression_small_cls": ("net_regression_small", dict(softmax_classes=4, spread_class_power=True, late_dropout_prob=0.0), "prob"),
    "regression_ladder_small": ("net_regression_ladder_small", dict(likelihood_loss=True), "lik"),
    "classification": ("net_classification", dict(softmax_classes=6, spread_class_power=True), "logits"),
    "classification_ladder": ("net_classification_ladder", dict(softmax_classes=5), "logits"),
}
FORK_NET_GRADS = ("encoder.initial_block.conv.weight", "decoder.scalar_decoder_1.layers.0.conv.bias", "decoder.scalar_decoder_2.layers.1.conv1x3_2.weight")


def fork_net_inputs():
    x = golden_input((1, 3, 64, 128), 91)
    tgt = golden_input((1, 64, 128), 92) * 2.0
    tgt[golden_input((1, 64, 128), 93) < 0.6] = -1.0
    return x, tgt, golden_labels((1, 64, 128), 94, n_classes=5)


def fork_net_loss(ev, kind, out, tgt, labels):
    """The criterion train/main.py:174-186 pairs with each head."""
    if kind == "reg":
        return ev.L1LossMasked()(out, tgt)
    if kind == "lik":
        return ev.LogLikelihoodLossMasked(opt_eps=0.05)(out[:, 0], out[:, 1].pow(2), tgt)
    if kind == "prob":
        return ev.L1LossClassProbMasked()(out[0], out[1], tgt)
    return ev.CrossEntropyLoss2d()(out[0], labels)


def fork_nets(path):
    """The five remaining train/net_*.py graphs (net_regression, _small, _ladder_small, net_classification, _ladder): state_dict layout,
    train-mode output, loss of the matching criterion and a few gradients at 1 x 3 x 64 x 128 (dropout p forced to 0: deterministic)."""
    import contextlib
    import importlib
    import io
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(REF, "train"))
    res = {}
    x, tgt, labels = fork_net_inputs()
    for name, (modname, kw, kind) in FORK_NETS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            ref_ev = importlib.import_module("evaluation_functions")
            m = importlib.import_module(modname).Net(**kw)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        res[name + "_keys"] = np.array([f"{k} {sh}" for k, sh in shapes.items()])
        sd = O.make_state_dict(shapes, seed=41)
        if "class_power" in sd:
            sd["class_power"] = torch.linspace(0.2, 1.8, sd["class_power"].numel()).view_as(sd["class_power"])
        m.load_state_dict(sd, strict=True)
        m.train()
        out = m(x)
        first = out[0] if isinstance(out, tuple) else out
        res[name + "_out"] = first.detach().numpy()
        loss = fork_net_loss(ref_ev, kind, out, tgt, labels)
        loss.backward()
        res[name + "_loss"] = np.float64(loss.item())
        for k in FORK_NET_GRADS + (("class_power",) if kind == "prob" else ()):
            gten = m.get_parameter(k).grad
            res[f"{name}_g_{k}"] = gten.numpy()
    np.savez_compressed(path, **res)


IMAGENET_GRADS = ("features.encoder.initial_block.conv.weight", "features.encoder.layers.10.bn2.weight", "classifier.linear.bias")


def imagenet_inputs(classes=1000):
    g = torch.Generator().manual_seed(3000)
    x = torch.randn(2, 3, 224, 224, generator=g)            # normalised-image-like
    y = torch.tensor([3, classes - 5])
    return x, y


def imagenet_net(path):
    """imagenet/erfnet_imagenet.ERFNet(1000) (BASELINE config 3): state_dict layout, train-mode logits (BatchNorm batch statistics,
    dropout p forced to 0), nn.CrossEntropyLoss, accuracy() top-1/top-5 of imagenet/main.py:288-301 (restated with reshape: its
    ``.view(-1)`` on the transposed prediction no longer runs on current torch), a few gradients, and parameters after two
    SGD(0.002, momentum 0.9, weight decay 1e-4) steps; plus the eval-mode logits after those steps."""
    import importlib
    sys.path.insert(0, os.path.join(REF, "imagenet"))
    ref = importlib.import_module("erfnet_imagenet")
    m = ref.ERFNet(1000)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    res = {"keys": np.array([f"{k} {sh}" for k, sh in shapes.items()])}
    m.load_state_dict(O.make_state_dict(shapes, seed=51), strict=True)
    x, y = imagenet_inputs()
    opt = torch.optim.SGD(m.parameters(), 0.002, momentum=0.9, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss()
    m.train()
    for step in range(2):
        out = m(x)
        loss = crit(out, y)
        opt.zero_grad()
        loss.backward()
        if step == 0:
            res["logits"], res["loss"] = out.detach().numpy(), np.float64(loss.item())
            _, pred = out.detach().topk(5, 1, True, True)
            correct = pred.t().eq(y.view(1, -1).expand_as(pred.t()))
            res["top1"] = np.float64(correct[:1].reshape(-1).float().sum(0) * 100.0 / 2)
            res["top5"] = np.float64(correct[:5].reshape(-1).float().sum(0) * 100.0 / 2)
            for k in IMAGENET_GRADS:
                res["g_" + k] = m.get_parameter(k).grad.numpy().copy()
        opt.step()
    res["loss2"] = np.float64(loss.item())
    for k in IMAGENET_GRADS:
        res["p2_" + k] = m.get_parameter(k).detach().numpy().copy()
    res["rm2"] = m.features.encoder.layers[10].bn2.running_mean.numpy().copy()
    m.eval()
    with torch.no_grad():
        res["eval_logits"] = m(x).numpy()
    # accuracy() / CrossEntropyLoss KAT on a logits matrix with a spread of target ranks (and an ignored row for the loss)
    lg, tg = topk_kat_inputs()
    _, pred = lg.topk(5, 1, True, True)
    correct = pred.t().eq(tg.view(1, -1).expand_as(pred.t()))
    res["kat_topk"] = np.array([float(correct[:k].reshape(-1).float().sum(0) * 100.0 / tg.numel()) for k in (1, 3, 5)])
    lg2 = lg.clone().requires_grad_(True)
    tg2 = tg.clone()
    tg2[4] = -100
    l2 = torch.nn.CrossEntropyLoss()(lg2, tg2)
    l2.backward()
    res["kat_loss"], res["kat_dlogits"] = np.float64(l2.item()), lg2.grad.numpy()
    np.savez_compressed(path, **res)


def topk_kat_inputs():
    lg = (golden_input((64, 37), 55) - 0.5) * 6.0
    g = torch.Generator().manual_seed(56)
    tg = torch.randint(0, 37, (64,), generator=g)
    order = lg.argsort(1, descending=True)
    for i in range(0, 64, 3):          # every third row: make the target one of the top-6 classes
        tg[i] = order[i, (i // 3) % 6]
    return lg, tg


def mean_teacher_inputs():
    x1 = golden_input((1, 3, 64, 128), 101)
    x2 = (x1 + 0.05 * (golden_input((1, 3, 64, 128), 102) - 0.5)).clamp(0, 1)     # the second augmentation of the same frame
    tgt = golden_input((1, 64, 128), 103) * 2.0
    tgt[golden_input((1, 64, 128), 104) < 0.6] = -1.0
    return x1, x2, tgt


def mean_teacher_step(student, teacher, ev, opt, x1, x2, tgt, consistency_weight, update_teacher):
    """One iteration of train/main.py:297-392 with --mean-teacher on a regression net (l.350-381): student on images1, teacher on images2,
    L1LossMasked on both, MSELossWeighted consistency, consistency.backward(retain_graph) + prediction.backward, optimizer.step, teacher update."""
    out_s = student(x1)
    opt.zero_grad()
    out_t = teacher(x2)
    loss_c = ev.MSELossWeighted(3.0)(out_s, out_t, consistency_weight)
    loss_s = ev.L1LossMasked()(out_s, tgt)
    loss_t = ev.L1LossMasked()(out_t, tgt)
    loss_c.backward(retain_graph=True)
    loss_s.backward()
    opt.step()
    update_teacher(0.01)                                    # DISCOUNT_RATE_START, train/main.py:48
    return float(loss_s.detach()), float(loss_t.detach()), float(loss_c.detach())


def mean_teacher(path):
    """Three mean-teacher iterations of the reference's net_regression_small.Net pair (student seed 41, teacher seed 43) with
    Adam(1e-4, (.9,.999), eps 1e-4, wd 1e-6) on the student (train/main.py:43-46, 212-213): the three losses per step and teacher / student
    parameter samples afterwards."""
    import contextlib
    import importlib
    import io
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(REF, "train"))
    with contextlib.redirect_stdout(io.StringIO()):
        ev = importlib.import_module("evaluation_functions")
        mod = importlib.import_module("net_regression_small")
        student, teacher = mod.Net(), mod.Net()
    shapes = {k: tuple(v.shape) for k, v in student.state_dict().items()}
    student.load_state_dict(O.make_state_dict(shapes, seed=41))
    teacher.load_state_dict(O.make_state_dict(shapes, seed=43))
    for target in teacher.parameters():                    # what copyWeightsToModelNoGrad effectively does (train/main.py:130-135)
        target.requires_grad = False
    x1, x2, tgt = mean_teacher_inputs()
    opt = torch.optim.Adam([p for p in student.parameters()], 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)

    def update(rate):                                      # copyWeightsToModelWithDiscount (train/main.py:137-139)
        for source, target in zip(student.parameters(), teacher.parameters()):
            target.data = target.data * (1 - rate) + rate * source.data

    student.train(), teacher.train()
    losses = [mean_teacher_step(student, teacher, ev, opt, x1, x2, tgt, 0.5, update) for _ in range(3)]
    res = {"losses": np.array(losses)}
    for k in ("decoder.scalar_output_conv.weight", "encoder.initial_block.conv.bias"):
        res["student_" + k] = student.get_parameter(k).detach().numpy().copy()
        res["teacher_" + k] = teacher.get_parameter(k).detach().numpy().copy()
    np.savez_compressed(path, **res)


def erfnet_nobn(path):
    """eval/erfnet_nobn.ERFNet(20) (the net eval_forwardTime.py times): state_dict layout and eval logits at 1 x 3 x 64 x 128."""
    import importlib
    sys.path.insert(0, os.path.join(REF, "eval"))
    ref = importlib.import_module("erfnet_nobn")
    m = ref.ERFNet(20)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(O.make_state_dict(shapes, seed=71), strict=True)
    m.eval()
    x = golden_input((1, 3, 64, 128), 141)
    with torch.no_grad():
        y = m(x)
        ye = m(x, only_encode=True)
    np.savez_compressed(path, keys=np.array([f"{k} {sh}" for k, sh in shapes.items()]), logits=y.numpy(), enc_logits=ye.numpy())


def fork_loss_inputs():
    """Deterministic operands of the fork-loss KATs (regenerated by the tests, not stored)."""
    b, c, h, w = 2, 6, 12, 20
    tgt = golden_input((b, h, w), 21) * 2.0
    tgt[golden_input((b, h, w), 22) < 0.6] = -1.0                      # ~60 % unlabelled pixels
    return dict(out=golden_input((b, 1, h, w), 23) * 2.0, var=golden_input((b, h, w), 24) + 0.05, target=tgt,
                prob=torch.softmax((golden_input((b, c, h, w), 25) - 0.5) * 4, 1), cost=golden_input((1, c, 1, 1), 26) * 2.0,
                cvar=golden_input((1, c, 1, 1), 27) + 0.1, teacher=golden_input((b, 1, h, w), 28) * 2.0,
                labels=golden_labels((b, h, w), 29, n_classes=c))


def fork_losses(path):
    """KATs of the fork's masked regression losses and accuracy metrics (train/evaluation_functions.py:32-163)."""
    import warnings
    sys.path.insert(0, os.path.join(REF, "train"))
    warnings.simplefilter("ignore")
    import evaluation_functions as ref_ev
    d = fork_loss_inputs()
    res = {}

    def run(name, crit, args, grad_idx):
        leaves = [a.clone().requires_grad_(i in grad_idx) for i, a in enumerate(args)]
        loss = crit(*leaves)
        loss.backward()
        res[name + "_loss"] = np.float64(loss.item())
        for i in grad_idx:
            res[f"{name}_g{i}"] = leaves[i].grad.numpy()

    run("l1", ref_ev.L1LossMasked(), (d["out"], d["target"]), (0,))
    run("loglik", ref_ev.LogLikelihoodLossMasked(opt_eps=1e-4), (d["out"][:, 0], d["var"], d["target"]), (0, 1))
    run("l1cp", ref_ev.L1LossClassProbMasked(), (d["prob"], d["cost"], d["target"]), (0, 1))
    run("loglikcp", ref_ev.LogLikelihoodLossClassProbMasked(opt_eps=1e-4), (d["prob"], d["cost"], d["cvar"], d["target"]), (0, 1, 2))
    leaves = [d["out"].clone().requires_grad_(True), d["teacher"].clone().requires_grad_(True)]
    loss = ref_ev.MSELossWeighted(3.0)(leaves[0], leaves[1], 0.25)
    loss.backward()
    res["mse_loss"], res["mse_g0"], res["mse_g1"] = np.float64(loss.item()), leaves[0].grad.numpy(), leaves[1].grad.numpy()
    acc = ref_ev.ClassificationAccuracy()(d["prob"], d["labels"])
    macc, cacc = ref_ev.MeanAccuracy(d["prob"].shape[1])(d["prob"], d["labels"])
    res["acc"], res["mean_acc"], res["class_acc"] = np.float64(acc.item()), np.float64(macc.item()), np.array([float(v) for v in cacc])
    np.savez_compressed(path, **res)


if __name__ == "__main__":
    if len(sys.argv) > 1:      # regenerate single fixtures: python oracle/make_golden.py fork_nets ...
        for fn in sys.argv[1:]:
            globals()[fn](os.path.join(OUT, fn + ".npz"))
    else:
        main()
