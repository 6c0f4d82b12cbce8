"""Shared body of the I/O-side and mean-teacher tests: uint8 ingest == torchvision ToTensor (u8 / 255) + layout, label LUT == ToLabel +
Relabel, Colorize, and the teacher update of train/main.py:140-142 -- each against the reference's own arithmetic restated inline in torch
(these are 1-3 line element-wise definitions; the reference classes need PIL images / are functions of train/main.py that cannot be imported)."""
import numpy as np
import torch

from erfnet_pytorch_b200 import erfnet as E
from erfnet_pytorch_b200 import ops, transform
from erfnet_pytorch_b200.optim import MeanTeacher
from oracle import erfnet_oracle as O


def check_ingest(device):
    g = torch.Generator().manual_seed(5)
    for shape in ((2, 16, 24, 3), (1, 5, 7, 3)):                       # 4-pixel groups, and a ragged tail
        img = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
        ref = img.permute(0, 3, 1, 2).float().div(255)                  # ToTensor
        got = ops.ingest_u8(img.to(device), 4, torch.float32)
        assert got.shape == shape[:3] + (4,)
        assert torch.equal(got[..., :3].cpu(), ref.permute(0, 2, 3, 1)) and float(got[..., 3].abs().max()) == 0.0   # bit-exact
        mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
        gotn = ops.ingest_u8(img.to(device), 4, torch.float32, mean.to(device), std.to(device))
        refn = (ref - mean[None, :, None, None]) / std[None, :, None, None]   # Normalize
        np.testing.assert_allclose(gotn[..., :3].cpu().numpy(), refn.permute(0, 2, 3, 1).numpy(), rtol=1e-6, atol=1e-7)
        gb = ops.ingest_u8(img.to(device), 4, torch.bfloat16)
        assert torch.equal(gb[..., :3].cpu(), ref.permute(0, 2, 3, 1).to(torch.bfloat16))


def check_model_accepts_uint8(device, precision="fp32"):
    """model(uint8 HWC batch) == model(ToTensor(batch)): bit for bit in fp32 check mode; in bf16 the float path's stem reads the fp32
    pixels while the ingest kernel stores bf16 pixels, so the two agree within the bf16 gate."""
    m = E.ERFNet(20)
    m.load_state_dict(O.make_state_dict(O.erfnet_param_shapes(20), 0))
    m.precision = precision
    m = m.to(device).eval()
    img = torch.randint(0, 256, (1, 32, 64, 3), generator=torch.Generator().manual_seed(6), dtype=torch.uint8)
    with torch.no_grad():
        a = m(img.to(device))
        b = m(img.permute(0, 3, 1, 2).float().div(255).to(device))
    if precision == "fp32":
        assert torch.equal(a, b)
    else:
        assert (a - b).abs().max().item() <= 2e-2 * b.abs().max().item()


def check_labels(device):
    g = torch.Generator().manual_seed(7)
    lab = torch.randint(0, 256, (2, 9, 13), generator=g, dtype=torch.uint8)
    lab[0, 0, :5] = 255
    out = transform.Relabel(255, 19)(lab.to(device))
    ref = lab.long()
    ref[ref == 255] = 19
    assert out.dtype == torch.int64 and torch.equal(out.cpu(), ref)
    out64 = transform.Relabel(255, 19)(lab.long().to(device))
    assert torch.equal(out64.cpu(), ref)
    tl = transform.ToLabel()(lab[0].to(device))
    assert tl.shape == (1, 9, 13) and torch.equal(tl.cpu()[0], lab[0].long())
    table = torch.arange(256).flip(0)
    assert torch.equal(transform.LabelTable(table)(lab.to(device)).cpu(), table[lab.long()])
    # Colorize (eval/transform.py:73-93), restated
    gray = torch.randint(0, 25, (1, 9, 13), generator=g)
    col = transform.Colorize(22)
    cmap = col.cmap
    ref_img = torch.zeros(3, 9, 13, dtype=torch.uint8)
    for label in range(0, len(cmap)):
        mask = gray[0] == label
        for ch in range(3):
            ref_img[ch][mask] = cmap[label][ch]
    assert torch.equal(col(gray.to(device)).cpu(), ref_img)
    assert torch.equal(col(gray.to(torch.uint8).to(device)).cpu(), ref_img)


def check_mean_teacher(device):
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Conv2d(5, 2, 1)).to(device)
    student, teacher = mk(), mk()
    s0 = [p.detach().clone() for p in student.parameters()]
    t0 = [p.detach().clone() for p in teacher.parameters()]
    mt = MeanTeacher(student, teacher)
    assert all(not p.requires_grad for p in teacher.parameters()) and all(p.requires_grad for p in student.parameters())
    for a, b in zip(student.parameters(), s0):
        assert torch.equal(a.detach(), b)                               # flattening keeps the values
    mt.update(0.25)
    for t, a, b in zip(teacher.parameters(), t0, s0):
        np.testing.assert_allclose(t.detach().cpu().numpy(), (a * (1 - 0.25) + 0.25 * b).cpu().numpy(), rtol=1e-6, atol=1e-7)
    with torch.no_grad():
        for p in student.parameters():
            p.add_(1.0)                                                 # in-place student updates are seen through the flat view
    mt.update(1.0)
    for t, b in zip(teacher.parameters(), s0):
        np.testing.assert_allclose(t.detach().cpu().numpy(), (b + 1.0).cpu().numpy(), rtol=1e-6)


def check_iou_one_hot(device):
    """eval/iouEval.py:37-46: [B,C,H,W] one-hot predictions / targets give the same counts as the label maps."""
    from erfnet_pytorch_b200.iouEval import iouEval
    g = torch.Generator().manual_seed(8)
    p = torch.randint(0, 20, (2, 1, 12, 16), generator=g)
    t = torch.randint(0, 20, (2, 1, 12, 16), generator=g)
    a, b = iouEval(20), iouEval(20)
    a.addBatch(p.to(device), t.to(device))
    hot = lambda v: torch.zeros(2, 20, 12, 16).scatter_(1, v, 1)
    b.addBatch(hot(p).to(device), hot(t).to(device))
    assert torch.equal(a.getIoU()[1], b.getIoU()[1]) and torch.equal(a.tp, b.tp) and torch.equal(a.fp, b.fp) and torch.equal(a.fn, b.fn)
    tp, fp, fn = O.iou_counts(p.numpy(), t.numpy(), 20, 19)
    assert np.array_equal(b.tp.numpy().astype(np.int64), tp) and np.array_equal(b.fn.numpy().astype(np.int64), fn)
