"""GPU: the ImageNet net and training step (BASELINE config 3), row-wise CE / top-k, flat SGD, the mean-teacher update and the I/O-side
kernels (uint8 ingest, label tables, colouring) through the C ABI."""
import pytest
import torch

from tests.imagenet_cases import build_imagenet, check_ce_topk, check_imagenet_flat_trainer, check_imagenet_net, check_pretrained_encoder_handoff
from tests.io_cases import check_ingest, check_iou_one_hot, check_labels, check_mean_teacher, check_model_accepts_uint8
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


def test_imagenet_net_fp32_golden(golden_dir):
    check_imagenet_net(golden_dir, "cuda")


def test_imagenet_flat_trainer(golden_dir):
    check_imagenet_flat_trainer(golden_dir, "cuda")


def test_pretrained_encoder_handoff():
    check_pretrained_encoder_handoff("cuda")


def test_ce_rows_topk(golden_dir):
    check_ce_topk(golden_dir, "cuda")


def test_imagenet_bf16_vs_fp32():
    """Batch 8 at 224 x 224, eval mode: bf16 storage vs the fp32 check mode of the same kernels."""
    x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(9)).cuda()
    outs = {}
    for precision in ("fp32", "bf16"):
        m = build_imagenet("cuda", precision).eval()
        with torch.no_grad():
            outs[precision] = m(x).float().cpu().numpy()
    assert outs["bf16"].shape == (8, 1000)
    assert_rel(outs["bf16"], outs["fp32"], 3e-2, "imagenet bf16 vs fp32")


def test_ingest_labels_colorize():
    check_ingest("cuda")
    check_labels("cuda")
    check_model_accepts_uint8("cuda", "fp32")
    check_model_accepts_uint8("cuda", "bf16")


def test_iou_one_hot_inputs():
    check_iou_one_hot("cuda")


def test_mean_teacher():
    check_mean_teacher("cuda")


def test_mean_teacher_training_vs_reference(golden_dir):
    from tests.mean_teacher_cases import check_mean_teacher_training
    check_mean_teacher_training(golden_dir, "cuda", "fp32")
